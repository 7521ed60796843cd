#!/usr/bin/env python
"""Micro-benchmark of the weight-gradient GEMMs of the update (reduction over the 2.0 M batch
rows) in the layouts torch / cuBLAS can be asked for.  TF32 on, CUDA events, median of 10."""
import json

import torch

torch.backends.cuda.matmul.allow_tf32 = True
M = 2_000_000


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    res = {}
    for name, (ka, kb) in {"dW2 [256 x M] x [M x 256]": (256, 256),
                           "dW3 [44 x M] x [M x 256]": (44, 256),
                           "dW1 [256 x M] x [M x 72]": (256, 72)}.items():
        A = torch.randn(M, ka, device="cuda")
        B = torch.randn(M, kb, device="cuda")
        ref = A.t().mm(B)
        v = {}
        v["A.t().mm(B)"] = timed(lambda: A.t().mm(B))
        v["B.t().mm(A).t()"] = timed(lambda: B.t().mm(A).t())
        v["einsum"] = timed(lambda: torch.einsum("mi,mj->ij", A, B))
        for chunks in (4, 16):
            Ac, Bc = A.view(chunks, M // chunks, ka), B.view(chunks, M // chunks, kb)
            v[f"bmm over {chunks} row chunks + sum"] = timed(
                lambda: torch.bmm(Ac.transpose(1, 2), Bc).sum(0))
            out = torch.bmm(Ac.transpose(1, 2), Bc).sum(0)
            assert torch.allclose(out, ref, rtol=1e-2, atol=1e-1 * float(ref.abs().max()))
        gb = (ka + kb) * M * 4 / 1e9
        res[name] = {"ms": v, "min_ms_at_hbm_peak": gb / 6576 * 1e3}
        del A, B
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
