#!/usr/bin/env python
"""Drop-in for the reference's scripts/build_resnet2d_embeddings.py (same flags, defaults, output files):
slice-mean ResNet2D embeddings for every manifest row -> resnet2d_<mh>_<ch>.parquet + .json sidecar.
Run under `torchrun --nproc-per-node N` to shard the subjects over N GPUs of one box."""
import argparse
import json
from pathlib import Path

import _bootstrap  # noqa: F401
from pd_fusion_b200.data.openneuro_features import _hash_config, _hash_file, build_resnet2d_embeddings
from pd_fusion_b200.parallel import init_distributed

FLAGS = [  # (flag, type, default, nargs)
    ("--backbone", str, "resnet18", None), ("--target-shape", int, [160, 160, 160], 3), ("--slice-axis", int, 2, None),
    ("--slice-count", int, 24, None), ("--input-size", int, 224, None), ("--batch-size", int, 32, None), ("--tta", int, 1, None),
    ("--max-rotation-deg", float, 5.0, None), ("--max-translation", float, 0.05, None), ("--intensity-scale", float, 0.1, None),
    ("--intensity-shift", float, 0.1, None), ("--noise-std", float, 0.01, None),
]


def main():
    ap = argparse.ArgumentParser(description="Build ResNet2D embeddings for ds001907")
    ap.add_argument("--manifest", type=str, required=True)
    ap.add_argument("--out-dir", type=str, default="data/processed/openneuro_ds001907/embeddings_resnet2d")
    for flag, typ, default, nargs in FLAGS:
        ap.add_argument(flag, type=typ, default=default, **({"nargs": nargs} if nargs else {}))
    a = ap.parse_args()
    rank, _, _ = init_distributed()
    manifest, out_dir = Path(a.manifest), Path(a.out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    # key order is irrelevant (the hash sorts), value reprs are part of the cache key
    cfg = {name.lstrip("-").replace("-", "_"): getattr(a, name.lstrip("-").replace("-", "_")) for name, *_ in FLAGS}
    df = build_resnet2d_embeddings(manifest, out_dir, cfg)
    if rank == 0:
        meta = out_dir / f"resnet2d_{_hash_file(manifest)}_{_hash_config(cfg)}.json"
        meta.write_text(json.dumps({"manifest": str(manifest), "config": cfg}, indent=2))
        print(f"Saved embeddings to {len(df)} rows in {out_dir}")


if __name__ == "__main__":
    main()
