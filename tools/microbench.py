#!/usr/bin/env python
"""Per-operation device timings (CUDA events, median of N) on the bench workload — used to fill
profiles/ and to decide what to optimise next.   python tools/microbench.py [--workload cfg3]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def timeit(torch, fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--pop", type=int, default=0)
    args = ap.parse_args()
    import torch
    from tracefly_b200 import _lib as L
    from tracefly_b200 import batched as B
    cfg = dict(bench.WORKLOADS[args.workload])
    if args.pop:
        cfg["P"] = args.pop
    env, params, pop, ga = bench.build_gpu_workload(cfg, 0, torch, B, L)
    P, W = cfg["P"], cfg["W"]
    out = {}
    fit, _ = B.fitness_batch(pop, None, env, params, want_collisions=False)
    nxt, nfit = torch.empty_like(pop), torch.empty_like(fit)
    out["fitness"] = timeit(torch, lambda: B.fitness_batch(pop, None, env, params, out=fit, want_collisions=False))
    for k in (2, 10, 1024):
        tk = B.TopK(P, k)
        idx = torch.empty(k, dtype=torch.int32, device="cuda")
        out[f"topk_{k}"] = timeit(torch, lambda: tk(fit, out=idx))
        out[f"topk_{k}_worst"] = timeit(torch, lambda: tk(fit, worst=True, out=idx))
    top = B.rank_topk(fit, 10)
    out["generation_fused"] = timeit(torch, lambda: B.generation(env, pop, fit, top, ga, params, seed=1, generation=1, out_pop=nxt, out_fit=nfit))
    out["ga_step_unfused"] = timeit(torch, lambda: B.ga_step(pop, fit, top, ga, seed=1, generation=1, out=nxt, want_parity=False))
    ga2 = L.default_ga_params(selection=L.SEL_TRUNCATION, n_top=10)
    out["generation_fused_truncation"] = timeit(torch, lambda: B.generation(env, pop, fit, top, ga2, params, seed=1, generation=1, out_pop=nxt, out_fit=nfit))
    # the same fused generation with no environment at all (segment arithmetic only): what the row traffic costs
    empty = B.Environment.empty()
    out["generation_fused_no_environment"] = timeit(torch, lambda: B.generation(empty, pop, fit, top, ga, params, seed=1, generation=1, out_pop=nxt, out_fit=nfit))
    out["generation_fused_no_environment_truncation"] = timeit(torch, lambda: B.generation(empty, pop, fit, top, ga2, params, seed=1, generation=1, out_pop=nxt, out_fit=nfit))
    out["fitness_no_environment"] = timeit(torch, lambda: B.fitness_batch(pop, None, empty, params, out=fit, want_collisions=False))
    perm = torch.randperm(P, device="cuda")
    out["gather_rows(torch.index_select, random permutation)"] = timeit(torch, lambda: torch.index_select(pop, 0, perm, out=nxt))
    out["fill_pop(torch.zero_, write only)"] = timeit(torch, lambda: nxt.zero_())
    out["sum_pop(torch.sum, read only)"] = timeit(torch, lambda: torch.sum(pop))
    src = torch.empty_like(pop)
    out["copy_pop_d2d(torch)"] = timeit(torch, lambda: src.copy_(pop))
    # ---- the north-star-named variants and the ones that stay on the LSU path (tf_stream_host.cuh: stream_eligible:
    #      S > 1 samples per segment, ragged lengths); their roofline fractions
    peak = bench.measured_peak_gbs()[0]
    gau = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=4, crossover=L.CX_UNIFORM, mutation=L.MUT_GAUSSIAN,
                              mutation_rate=0.1, sigma=1.5, n_elite=2)
    out["generation_fused_uniform_crossover_gaussian_mutation(TMA path, both parent rows staged)"] = timeit(
        torch, lambda: B.generation(env, pop, fit, top, gau, params, seed=1, generation=1, out_pop=nxt, out_fit=nfit))
    gam = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=4, mutation=L.MUT_GAUSSIAN, mutation_rate=0.1, sigma=1.5, n_elite=2)
    out["generation_fused_single_point_gaussian_mutation(TMA path)"] = timeit(
        torch, lambda: B.generation(env, pop, fit, top, gam, params, seed=1, generation=1, out_pop=nxt, out_fit=nfit))
    p3 = L.Params.from_buffer_copy(params)
    p3.samples_per_segment = 3
    out["fitness_S3_samples_per_segment(TMA path)"] = timeit(torch, lambda: B.fitness_batch(pop, None, env, p3, out=fit, want_collisions=False), n=10)
    lens = torch.full((P,), W, dtype=torch.int32, device="cuda")
    lens[::3] = W - 7
    out["fitness_ragged_lengths(LSU path)"] = timeit(torch, lambda: B.fitness_batch(pop, lens, env, params, out=fit, want_collisions=False))
    B.fitness_batch(pop, None, env, params, out=fit, want_collisions=False)
    bytes_gen = (32 * W + 12) * P
    rep = {k: {"median_ms": v[0], "min_ms": v[1]} for k, v in out.items()}
    rep["generation_fused"]["algorithmic_GBs"] = bytes_gen / (out["generation_fused"][0] * 1e-3) / 1e9
    rep["fitness"]["algorithmic_GBs"] = (16 * W + 4) * P / (out["fitness"][0] * 1e-3) / 1e9
    for k in rep:
        if k.startswith("generation_fused"):
            rep[k]["frac_of_measured_hbm"] = bytes_gen / (out[k][0] * 1e-3) / 1e9 / peak
        if k.startswith("fitness"):
            rep[k]["frac_of_measured_hbm"] = (16 * W + 4) * P / (out[k][0] * 1e-3) / 1e9 / peak
    rep["copy_pop_d2d(torch)"]["GBs"] = 2 * 16 * W * P / (out["copy_pop_d2d(torch)"][0] * 1e-3) / 1e9
    print(json.dumps({"workload": args.workload, "P": P, "W": W, "ops": rep}, indent=1))


if __name__ == "__main__":
    main()
