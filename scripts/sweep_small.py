"""Shape sweep of the per-modality models only (D = 64 / 128), fused chain vs per-layer launches (MMAD_NO_SMALLNET=1)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

r = bench.bench_shapes(0, torch.device("cuda:0"), rows=int(os.environ.get("ROWS", 4 * 148 * 128)), steps=20)
print(json.dumps({k: {p: round(v["samples_per_s"] / 1e6, 1) for p, v in rec.items()} for k, rec in r.items() if k.startswith(("D64", "D128"))}))
