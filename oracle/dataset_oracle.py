"""CPU restatement (numpy / torch, test infrastructure only) of the feature handling in the reference's MomentDataset.__getitem__
(hirest_dataset.py:323-407): linspace subsample / repeat-pad to n_model_frames and ASR sentence-feature warping.  Pinned against the
reference class itself by tests/golden/dataset.pt (oracle/make_golden_dataset.py)."""
import numpy as np
import torch


def resample(features: torch.Tensor, n_model_frames: int) -> torch.Tensor:
    """hirest_dataset.py:333-356 (video) / :383-403 (warped ASR)."""
    if n_model_frames <= 0:
        return features
    n = features.shape[0]
    if n > n_model_frames:
        ids = np.linspace(0, n - 1, n_model_frames).astype(int)
        return features[torch.from_numpy(ids)]
    x = torch.zeros((n_model_frames, features.shape[1]))
    j = 0
    for k in range(n):
        for _ in range((k * n_model_frames) // n, ((k + 1) * n_model_frames) // n):
            x[j] = features[k]
            j += 1
    return x


def warp_asr(asr_features: torch.Tensor, subs, len_vid: int) -> torch.Tensor:
    """hirest_dataset.py:370-381: subs = [(start_seconds, end_seconds)], one per feature row, applied in order."""
    out = torch.zeros(len_vid, asr_features.shape[1]).float()
    for i, (start, end) in enumerate(subs):
        out[start:end] = asr_features[i]
    return out
