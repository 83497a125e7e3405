"""cython3dmodelrenderer_b200 -- the Version C rendering hot path of oKatanaaa/Cython3DModelRenderer
(`AdvancedPixelBufferFiller.render_model`) as hand-written sm_100a CUDA behind a C ABI.

    from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller   # instead of crender.cy.pixel_buffer_filler

    from cython3dmodelrenderer_b200 import Model                       # optionally, instead of crender.cy.data_structures

    from cython3dmodelrenderer_b200 import Renderer, GuroIllumination  # optionally, instead of crender.cy / crender.cy.illumination:
                                                                       # the illumination runs on the device buffers

Everything else of the reference (iterators, run.py) is used unchanged; upstream's own Renderer and GuroIllumination
also drive this filler as they are (live NumPy views).
"""
from ._lib import CrenderError, build, load_library, projection_matrix  # noqa: F401
from .pixel_buffer_filler import AdvancedPixelBufferFiller  # noqa: F401
from .pipeline import HostFramePipeline, HostImagePipeline  # noqa: F401
from .model import Model  # noqa: F401   (drop-in for crender.cy.data_structures.Model)
from .illumination import GuroIllumination, IlluminationDrawer, NoIllumination  # noqa: F401   
from .renderer import Renderer  # noqa: F401

__all__ = ["AdvancedPixelBufferFiller", "HostFramePipeline", "HostImagePipeline", "Model", "Renderer", "GuroIllumination",
           "IlluminationDrawer", "NoIllumination", "CrenderError", "build", "load_library", "projection_matrix"]
