"""pd_fusion_b200 -- B200-native drop-in for the imaging-embedding + fusion hot path of
Ardbiu/robust-multimodal-pd (`pd_fusion`).  Sub-packages mirror the reference's module paths
(`data.openneuro_features`, `models.*`, `evaluation.evaluate`); all numerics run in
libpdfusion_b200.so (hand-written sm_100a CUDA), loaded through ctypes in `_lib`."""
__version__ = "0.1.0"
