"""hirest_b200 — B200-native (sm_100a) implementation of the HiREST video-text inference hot path.

Public surface (mirrors the reference's names for this path):
  hirest_b200.eva_clip.EVA_CLIP / build_eva_model_and_transforms   (EVA_clip/eva_model.py, eva_clip.py)
  hirest_b200.retrieval                                            (inference_video_retrieval.py scoring)
The compute lives in libhirest_b200.so (C ABI: include/hirest_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
