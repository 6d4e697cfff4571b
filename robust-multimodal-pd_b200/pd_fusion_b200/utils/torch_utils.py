"""Device selection (reference: utils/torch_utils.py:4-12).  The B200 build pins cuda:<LOCAL_RANK>."""
import os

import torch


def get_torch_device(prefer_mps: bool = True) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("pd_fusion_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
