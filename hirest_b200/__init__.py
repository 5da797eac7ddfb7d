"""hirest_b200 — B200-native (sm_100a) implementation of the HiREST video-text inference hot path.

Public surface (mirrors the reference's names for this path):
  hirest_b200.eva_clip.EVA_CLIP / build_eva_model_and_transforms   (EVA_clip/eva_model.py, eva_clip.py)
  hirest_b200.retrieval                                            (inference_video_retrieval.py scoring)
The compute lives in libhirest_b200.so (C ABI: include/hirest_b200.h); there is no CPU fallback.
"""
__version__ = "0.2.0"


def install_as_eva_clip() -> None:
    """Make ``from eva_clip import build_eva_model_and_transforms`` resolve to this package, so the reference's callers run
    UNCHANGED: ``modeling.py:115-117`` and ``inference_video_retrieval.py:176-183`` do ``sys.path.append("./EVA_clip")`` followed by
    that import; a module already present in ``sys.modules`` wins over the path search.  Call once before importing them
    (e.g. ``python -c "import hirest_b200; hirest_b200.install_as_eva_clip(); import runpy; runpy.run_path('run.py', run_name='__main__')"``)."""
    import sys

    from . import eva_clip

    sys.modules["eva_clip"] = eva_clip
