"""cython3dmodelrenderer_b200 -- the Version C rendering hot path of oKatanaaa/Cython3DModelRenderer
(`AdvancedPixelBufferFiller.render_model`) as hand-written sm_100a CUDA behind a C ABI.

    from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller   # instead of crender.cy.pixel_buffer_filler

Everything else of the reference (Model, Renderer, illumination, run.py) is used unchanged.
"""
from ._lib import CrenderError, build, load_library, projection_matrix  # noqa: F401
from .pixel_buffer_filler import AdvancedPixelBufferFiller  # noqa: F401
from .pipeline import HostFramePipeline  # noqa: F401

__all__ = ["AdvancedPixelBufferFiller", "HostFramePipeline", "CrenderError", "build", "load_library", "projection_matrix"]
