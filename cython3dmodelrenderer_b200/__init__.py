"""cython3dmodelrenderer_b200 -- the Version C rendering hot path of oKatanaaa/Cython3DModelRenderer
(`AdvancedPixelBufferFiller.render_model`) as hand-written sm_100a CUDA behind a C ABI.

    from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller   # instead of crender.cy.pixel_buffer_filler

    from cython3dmodelrenderer_b200 import Model                       # optionally, instead of crender.cy.data_structures

Everything else of the reference (Renderer, illumination, run.py) is used unchanged.
"""
from ._lib import CrenderError, build, load_library, projection_matrix  # noqa: F401
from .pixel_buffer_filler import AdvancedPixelBufferFiller  # noqa: F401
from .pipeline import HostFramePipeline, HostImagePipeline  # noqa: F401
from .model import Model  # noqa: F401   (drop-in for crender.cy.data_structures.Model)

__all__ = ["AdvancedPixelBufferFiller", "HostFramePipeline", "HostImagePipeline", "Model", "CrenderError", "build", "load_library", "projection_matrix"]
