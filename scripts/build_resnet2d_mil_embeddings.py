#!/usr/bin/env python
"""Drop-in for the reference's scripts/build_resnet2d_mil_embeddings.py (same flags, defaults, output files):
per-slice (MIL bag) ResNet2D embeddings -> resnet2d_mil_<mh>_<ch>.npz + .json sidecar.
Run under `torchrun --nproc-per-node N` to shard the subjects over N GPUs of one box."""
import argparse
import json
from pathlib import Path

import _bootstrap  # noqa: F401
from pd_fusion_b200.data.openneuro_features import _hash_config, _hash_file, build_resnet2d_mil_embeddings
from pd_fusion_b200.parallel import init_distributed


def main():
    ap = argparse.ArgumentParser(description="Build ResNet2D MIL embeddings for ds001907")
    ap.add_argument("--manifest", type=str, required=True)
    ap.add_argument("--out-dir", type=str, default="data/processed/openneuro_ds001907/embeddings_resnet2d")
    ap.add_argument("--backbone", type=str, default="resnet50")
    ap.add_argument("--target-shape", type=int, nargs=3, default=[160, 160, 160])
    ap.add_argument("--slice-axis", type=int, default=2)
    ap.add_argument("--slice-axes", type=int, nargs="+", default=None)
    ap.add_argument("--slice-count", type=int, default=48)
    ap.add_argument("--slice-counts", type=int, nargs="+", default=None)
    ap.add_argument("--input-size", type=int, default=224)
    ap.add_argument("--batch-size", type=int, default=16)
    ap.add_argument("--tta", type=int, default=1)
    ap.add_argument("--max-rotation-deg", type=float, default=5.0)
    ap.add_argument("--max-translation", type=float, default=0.05)
    ap.add_argument("--intensity-scale", type=float, default=0.1)
    ap.add_argument("--intensity-shift", type=float, default=0.1)
    ap.add_argument("--noise-std", type=float, default=0.01)
    a = ap.parse_args()
    rank, _, _ = init_distributed()
    manifest, out_dir = Path(a.manifest), Path(a.out_dir)
    axes = a.slice_axes if a.slice_axes else [a.slice_axis]
    if a.slice_counts and len(a.slice_counts) != len(axes):
        raise ValueError("slice-counts must match length of slice-axes")
    counts = a.slice_counts if a.slice_counts else [a.slice_count] * len(axes)
    cfg = {k: getattr(a, k) for k in ("backbone", "target_shape", "input_size", "batch_size", "tta", "max_rotation_deg",
                                      "max_translation", "intensity_scale", "intensity_shift", "noise_std")}
    if len(axes) == 1:
        cfg["slice_axis"], cfg["slice_count"] = axes[0], counts[0]
    else:
        cfg["slice_axes"], cfg["slice_counts"] = axes, counts
    out_path = build_resnet2d_mil_embeddings(manifest, out_dir, cfg, axes, counts)
    if rank == 0:
        meta = out_dir / f"resnet2d_mil_{_hash_file(manifest)}_{_hash_config(cfg)}.json"
        meta.write_text(json.dumps({"manifest": str(manifest), "config": cfg}, indent=2))
        print(f"Saved MIL embeddings to {out_path}")


if __name__ == "__main__":
    main()
