"""Drop-in illumination drawers for fillers of this package.

Mirror crender/cy/illumination/illumination_drawer.py:5-13 and guro_illumination.py:6-27: same classes, same constructor,
same `draw_illumination(color_buffer, n_buffer)` that lights the colour buffer in place.  Upstream's GuroIllumination is five
NumPy expressions over the whole frame -- 30-60 ms at 1024^2, ten times what its own Version C rasterizer takes and eighty
times a frame of this package, so a caller who only swaps the filler still spends the frame in NumPy.  This
class recognises the live views of an `AdvancedPixelBufferFiller` of this package and runs the same arithmetic on the
device buffers they mirror (crb_guro: float32, np.sum's +0.0 identity, left-to-right sums -- bit-equal to upstream, pinned
by the reference-made `*_guro` goldens), then refreshes the colour view, so the array the caller passed in holds the lit
colours exactly as after upstream's in-place `color_buffer *= shadow_coeff`.

There is no NumPy path here: arrays that are not the live views of one of this package's fillers raise TypeError
(upstream's own class serves those).
"""
import numpy as np

from .pixel_buffer_filler import AdvancedPixelBufferFiller


class IlluminationDrawer:
    """illumination_drawer.py:5-8"""

    def draw_illumination(self, color_buffer, n_buffer):
        pass


class NoIllumination(IlluminationDrawer):
    """illumination_drawer.py:11-13"""

    def draw_illumination(self, color_buffer, n_buffer):
        pass


# noinspection PyDefaultArgument
class GuroIllumination(IlluminationDrawer):
    def __init__(self, light_direction=[0, 0, 1]):
        """guro_illumination.py:7-18: the direction the light falls in; kept negated and normalised, in float32."""
        light_direction = -np.asarray(light_direction, dtype='float32')
        self.light_direction = light_direction / np.linalg.norm(light_direction)

    def draw_illumination(self, color_buffer, n_buffer):
        """guro_illumination.py:20-27 on the device buffers behind the two live views; `color_buffer` is updated in place."""
        filler = AdvancedPixelBufferFiller.owner_of_views(color_buffer, n_buffer)
        if filler is None:
            raise TypeError("cython3dmodelrenderer_b200.GuroIllumination lights the buffers of a cython3dmodelrenderer_b200 "
                            "AdvancedPixelBufferFiller: pass the arrays its get_color_buffer() / get_normals_buffer() return "
                            "(there is no NumPy path here; crender.cy.illumination.GuroIllumination serves other arrays)")
        filler.illuminate_guro(self.light_direction)     # uploads what the caller wrote through the views first (live-view contract)
        filler.get_color_buffer()                        # the same array object, refreshed from the device
