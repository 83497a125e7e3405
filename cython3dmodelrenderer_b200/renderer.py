"""Drop-in `Renderer` (crender/cy/renderer.py:9-52): same constructor, same `render(model, normalize_model, random_colors)`
returning the filler's live colour buffer, same `reset_buffers()`.

Upstream's `render` is `render_model`, `draw_illumination(get_color_buffer(), get_normals_buffer())`, `get_color_buffer()`
(renderer.py:47-49).  With a filler and an illumination of this package the same three steps happen on the device and only
the lit colour buffer crosses PCIe (12 of the 28 bytes per pixel, once): the rasterizer, then crb_guro on the device buffers
(illumination.py), then one download.  Any other filler / illumination pair takes upstream's sequence verbatim (this class
does no arithmetic of its own either way).
"""
from .illumination import GuroIllumination, NoIllumination
from .pixel_buffer_filler import AdvancedPixelBufferFiller


class Renderer:
    def __init__(self, pixel_buffer_filler, illumination, triangle_iterator_type=None,
                 image_height=512, image_width=512, use_tqdm=True):
        """renderer.py:10-19 (the iterator type, the image size and `use_tqdm` are stored and unused there as well)."""
        self.pixel_buffer_filler = pixel_buffer_filler
        self.illumination = illumination
        self.triangle_iterator_type = triangle_iterator_type
        self.im_h = image_height
        self.im_w = image_width
        self.use_tqdm = use_tqdm

    def render(self, model, normalize_model=False, random_colors=True):
        """renderer.py:21-49.  Returns the filler's colour buffer (live float32 [h,w,3] view)."""
        if normalize_model:     # renderer.py:41-46, on the caller's model as upstream
            image_center = (self.im_h // 2, self.im_w // 2)
            image_span = min(image_center)
            model.scale(image_span / model.get_max_span())
            model.shift(- model.get_mean_vertex() + [image_center[0], image_center[1], -image_span])
        filler = self.pixel_buffer_filler
        on_device = isinstance(filler, AdvancedPixelBufferFiller)
        if on_device and type(self.illumination) is GuroIllumination:
            filler.render_model(model, prefetch=False)                      # (the colour worth downloading is the lit one)
            filler.illuminate_guro(self.illumination.light_direction)       # device buffers, nothing has crossed PCIe yet
        elif on_device and type(self.illumination) is NoIllumination:
            filler.render_model(model)
        else:
            filler.render_model(model)
            self.illumination.draw_illumination(filler.get_color_buffer(), filler.get_normals_buffer())
        return filler.get_color_buffer()

    def reset_buffers(self):
        """renderer.py:51-52"""
        pass
