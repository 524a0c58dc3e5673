#!/usr/bin/env python3
"""bench.py — the driver's benchmark contract for the Plonky3 hot path on B200.

Default workload (N=1) = BASELINE.json configs[1]: coset_lde_batch, KoalaBear, 2^20 rows x 100 cols, blowup 2
(added_bits = 1, shift = GENERATOR), output in the committed (bit-reversed-row) layout.  metric = NTT Gelem/s of LDE
output.  A "step" is one LDE of one synthetic matrix (uniform field elements, seed 1).

  value       device-resident throughput: input already in HBM, CUDA events on the launching stream, K steps.
              Input (419 MB) + output (839 MB) exceed the 126 MB L2, so no L2 flush is needed between iterations.
  e2e         the same metric through the reference-facing C-ABI call p3gpu_coset_lde_batch with HOST (pinned) buffers:
              H2D of the input and D2H of the result are inside the timed region.
  roofline    HBM roofline of the NTT pass kernel: algorithmic bytes of one LDE (read input once + write output once,
              SURVEY.md §8d: 1,258,291,200 B) / device time of the step (all launches of a step are the same kernel).
  cpu_baseline the oracle port (OpenMP C restatement, oracle/p3_oracle.c) on the host cores, bounded sample.
  others      (N=1 only) the remaining single-GPU BASELINE configs timed the same way: config 3 (Poseidon2 Merkle 2^22 x 100),
              config 4 (BabyBear 2^22 x 300 LDE + Keccak Merkle, FRI commit phase 2^23) and the config 5 hot path
              (KoalaBear 2^20 x 1312 trace commit + quotient commit + FRI commit phase, all device resident).

--impl reference: times the reference's CPU algorithm (the oracle port — the reference is Rust and cannot be built in this
image) on the same metric; rank 0 only under torchrun.
Multi-GPU (--gpus N under torchrun): the path shards by independent matrices/column blocks with no data-path collective
(SURVEY.md §8e); every rank runs the same per-GPU workload (weak scaling), timing = max over ranks.
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LOG_H, W, ADDED_BITS = 20, 100, 1
ALG_BYTES = ((1 << LOG_H) * W + (1 << (LOG_H + ADDED_BITS)) * W) * 4       # 1,258,291,200
OUT_ELEMS = (1 << (LOG_H + ADDED_BITS)) * W                                   # 209,715,200


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-others", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock + throttle reasons DURING the timed region (NVML, ~1 kHz; falls back to nvidia-smi)."""

    def __init__(self, idx=0):
        self.idx, self.samples, self.reasons, self.stop = idx, [], set(), False
        self.max_mhz = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
            while not self.stop:
                self.samples.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                try:
                    r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, b in bits.items():
                    if r & b:
                        self.reasons.add(n)
                time.sleep(0.001)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass

    def __enter__(self): self.t.start(); time.sleep(0.05); self.samples.clear(); return self
    def __exit__(self, *a): self.stop = True; self.t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------ CPU (oracle) legs
def _omp_setup():
    """One OpenMP thread per PHYSICAL core, spread over the sockets: measured on the bench host (2 x Xeon 8562Y+, 64 cores /
    128 threads) this is 2x faster than the default 128 threads (0.95 s vs 1.9-2.1 s per 2^20 x 100 LDE)."""
    try:
        import psutil
        phys = psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        phys = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(phys))
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    return int(os.environ["OMP_NUM_THREADS"])


def cpu_lde_throughput(budget_s=15.0):
    """Oracle port of coset_lde_batch on a bounded column sample of the same workload.  Returns (Gelem/s, cores, sample)."""
    cores = _omp_setup()
    from oracle import p3_oracle as O
    O.build(native=True)          # rebuild with -march=native for THIS host
    f = 1
    m = O.random_matrix(f, 1 << LOG_H, 4, seed=1)
    t0 = time.time(); O.coset_lde_batch(f, m, ADDED_BITS, O.generator(f)); t4 = time.time() - t0
    cols = int(max(4, min(W, 4 * budget_s / max(t4, 1e-3))))
    cols -= cols % 4
    m = O.random_matrix(f, 1 << LOG_H, cols, seed=1)
    t0 = time.time(); out = O.coset_lde_batch(f, m, ADDED_BITS, O.generator(f)); dt = time.time() - t0
    return out.size / dt / 1e9, cores, f"coset_lde_batch KoalaBear 2^{LOG_H} x {cols} of {W} cols, blowup 2, {dt:.1f} s, OpenMP {cores} threads"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _omp_setup()
    from oracle import p3_oracle as O
    O.build(native=True)
    f = 1
    cols = W if cores >= 16 else 16          # full workload on a many-core host, a 16-column sample on small hosts
    m = O.random_matrix(f, 1 << LOG_H, cols, seed=1)
    for _ in range(min(args.warmup, 1)):
        O.coset_lde_batch(f, m, ADDED_BITS, O.generator(f))
    steps = max(1, min(args.steps, 5))
    t0 = time.time()
    for _ in range(steps):
        out = O.coset_lde_batch(f, m, ADDED_BITS, O.generator(f))
    dt = (time.time() - t0) / steps
    v = out.size / dt / 1e9
    sample = f"coset_lde_batch KoalaBear 2^{LOG_H} x {cols} of {W} cols per step (bounded sample), OpenMP {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "coset_lde_batch output Gelem/s (KoalaBear 2^20 x 100, blowup 2)", "value": v, "unit": "Gelem/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3 * (W / cols),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (31-bit Montgomery)", "data": "synthetic",
        "config": {"workload": "coset_lde_batch KoalaBear 2^20 x 100, added_bits 1, shift GENERATOR (BASELINE configs[1])"},
        "cpu_baseline": {"value": v, "unit": "Gelem/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Rust (no toolchain in this image): this is the OpenMP C restatement oracle/p3_oracle.c; ms_per_step is scaled to 100 columns",
    }))


# ------------------------------------------------------------------------------------------------ GPU legs
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from plonky3_b200 import _lib
    from plonky3_b200.field import KoalaBear as KB, BabyBear as BB
    from plonky3_b200.gpu import Gpu
    from plonky3_b200.poseidon2 import default_poseidon2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    gpu = Gpu(local)
    for f in (KB, BB):
        for w in (16, 24):
            default_poseidon2(f, w).upload(gpu)
    dev = f"cuda:{local}"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """W warm-up steps, then K timed steps bracketed by barrier+synchronize, CUDA events, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = gpu.launches
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, (gpu.launches - l0)

    # ---- primary: coset LDE, device resident (weak scaling: one matrix per rank)
    g = torch.Generator(device=dev); g.manual_seed(1 + rank)
    x = torch.randint(0, KB.P, (1 << LOG_H, W), device=dev, dtype=torch.int32, generator=g)
    out = torch.empty((1 << (LOG_H + ADDED_BITS), W), device=dev, dtype=torch.int32)
    gpu._use_torch_stream()

    def lde_step():
        _lib.check(gpu.L.p3gpu_coset_lde_batch_dev(gpu.h, KB.id, x.data_ptr(), 1 << LOG_H, W, ADDED_BITS, KB.generator, out.data_ptr(), 1))

    with ClockSampler(local) as clk:
        ms, launches = timed(lde_step, args.steps, max(args.warmup, 3))
    value = world * OUT_ELEMS / (ms * 1e-3) / 1e9

    # ---- e2e: host-pointer C-ABI call, pinned host buffers, copies inside the timed region.
    # Every step = one p3gpu_coset_lde_batch call (H2D 419 MB -> LDE -> D2H 839 MB).  Two calls are kept in flight from two
    # host threads, each with its own libp3gpu context (own stream, scratch and twiddle cache) — the reference's DFT objects
    # are Clone + Sync and may be called concurrently the same way — so one call's D2H overlaps the other's H2D + compute
    # (full-duplex PCIe).  The strictly serial single-call latency is reported next to it.
    import threading as _th
    INFLIGHT = 2
    lanes = []
    for _ in range(INFLIGHT):
        lg = Gpu(local)
        hx = torch.empty((1 << LOG_H, W), dtype=torch.int32).pin_memory(); hx.copy_(x.cpu())
        hout = torch.empty((1 << (LOG_H + ADDED_BITS), W), dtype=torch.int32).pin_memory()
        lanes.append((lg, hx, hout))

    def e2e_call(lane):
        lg, hx, hout = lane
        _lib.check(lg.L.p3gpu_coset_lde_batch(lg.h, KB.id, hx.data_ptr(), 1 << LOG_H, W, ADDED_BITS, KB.generator, hout.data_ptr(), 1))

    def e2e_run(n_steps, inflight):
        def worker(lane, n):
            for _ in range(n):
                e2e_call(lane)
        ths = [_th.Thread(target=worker, args=(lanes[i], n_steps // inflight + (1 if i < n_steps % inflight else 0))) for i in range(inflight)]
        barrier()
        t0 = time.perf_counter()
        for t_ in ths: t_.start()
        for t_ in ths: t_.join()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if world > 1:
            tt = torch.tensor([dt], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); dt = float(tt.item())
        return dt / n_steps

    e2e_steps = max(4, min(args.steps, 10))
    for lane in lanes:
        e2e_call(lane)                                   # warm-up (allocations, twiddle tables)
    single_ms = e2e_run(max(2, e2e_steps // 2), 1)
    e2e_ms = e2e_run(e2e_steps, INFLIGHT)
    for lane in lanes:
        assert torch.equal(lane[2].to(dev), out), "e2e result differs from device-resident result"
    e2e = {"value": world * OUT_ELEMS / (e2e_ms * 1e-3) / 1e9, "unit": "Gelem/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": lanes[0][1].numel() * 4, "d2h_bytes_per_step": lanes[0][2].numel() * 4,
           "api": "p3gpu_coset_lde_batch (host pointers, pinned), 2 calls in flight from 2 host threads / 2 contexts",
           "single_call_ms": single_ms, "single_call_value": world * OUT_ELEMS / (single_ms * 1e-3) / 1e9, "timer": "host wall clock around the calls (device work is synchronous inside the call)"}
    del lanes

    line = {
        "metric": "coset_lde_batch output Gelem/s (KoalaBear 2^20 x 100, blowup 2)", "value": value, "unit": "Gelem/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 (31-bit Montgomery)", "data": "synthetic",
        "config": {"workload": "coset_lde_batch KoalaBear 2^20 x 100, added_bits 1, shift GENERATOR, bit-reversed rows (BASELINE configs[1])",
                   "per_gpu_matrices": 1, "l2_policy": "inputs+outputs (1.26 GB) exceed the 126 MB L2; no flush needed"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary(),
    }

    # ---- roofline of the dominant kernel (ntt_pass_kernel): every launch of the step is this kernel
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = json.loads(peaks_path.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    achieved = ALG_BYTES / (ms * 1e-3) / 1e9
    traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one LDE step (ncu --set full)
    tp = ROOT / "profiles" / "ncu_traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get("lde_step_dram_bytes")
    line["roofline"] = {"bound": "hbm", "kernel": "ntt_pass_pipe_kernel (all launches of an LDE step)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "algorithmic_bytes_per_step": ALG_BYTES, "launches_per_step": launches / args.steps,
                        "note": "integer-pipe bound, not HBM bound: 3.146e9 butterflies x (IMAD.HI + 2 IMAD + 4 ALU) = 73 % of the issued instructions; register-only butterfly loop peaks at 12.85/clk/SM = 0.86 ms floor (DESIGN.md 4.1)"}

    # ---- secondary workloads (single GPU only)
    if world == 1 and not args.no_others:
        others = {}
        k = max(2, min(args.steps, 5))
        # config 1: Radix2DitParallel forward NTT, BabyBear, 2^16 x 1 (parity case; device time of the single-column transform)
        x1 = torch.randint(0, BB.P, (1 << 16, 1), device=dev, dtype=torch.int32, generator=g)
        t, nl = timed(lambda: gpu.dft_batch(BB.id, _lib.DFT, x1), 20, 3)
        others["config1_dft_bb_2^16x1"] = {"us": t * 1e3, "launches": nl / 20}
        # config 3: MerkleTreeMmcs commit 2^22 x 100 KoalaBear, Poseidon2-16 sponge, cap 0
        xm = torch.randint(0, KB.P, (1 << 22, 100), device=dev, dtype=torch.int32, generator=g)
        t, nl = timed(lambda: gpu.merkle_commit(KB.id, _lib.HASH_POSEIDON2_W16, [xm]), k, 1)
        others["merkle_commit_poseidon2_w16_kb_2^22x100"] = {"ms": t, "Mleaf_per_s": (1 << 22) / t / 1e3, "Mperm_per_s": 58720255 / t / 1e3,
                                                            "alg_GBps": 1.946e9 / (t * 1e-3) / 1e9, "launches": nl / k}
        del xm
        # config 5 leaf shape: Poseidon2-24 sponge + Poseidon2-16 compress over a 2^21 x 328 slice (quarter of 1312 columns)
        xw = torch.randint(0, KB.P, (1 << 21, 328), device=dev, dtype=torch.int32, generator=g)
        t, nl = timed(lambda: gpu.merkle_commit(KB.id, _lib.HASH_POSEIDON2_W24, [xw]), k, 1)
        others["merkle_commit_poseidon2_w24_kb_2^21x328"] = {"ms": t, "Mperm_per_s": ((1 << 21) * 21 + (1 << 21) - 1) / t / 1e3}
        del xw
        # config 4 (full size): BabyBear 2^22 x 300, TwoAdicFriPcs::commit = LDE blowup 2 + Keccak Merkle, then the FRI
        # commit phase on a 2^23 EF4 codeword with fixed betas (arities [3]*7+[1], cap_height 3)
        betas = np.random.default_rng(2).integers(0, BB.P, size=(10, 4), dtype=np.uint32)
        xb = torch.randint(0, BB.P, (1 << 22, 300), device=dev, dtype=torch.int32, generator=g)
        t, nl = timed(lambda: gpu.pcs_commit(BB.id, _lib.HASH_KECCAK, xb, 1), 2, 1)
        others["config4_pcs_commit_keccak_bb_2^22x300"] = {"ms": t, "launches": nl / 2, "lde_out_Gelem_per_s": (1 << 23) * 300 / t / 1e6}
        del xb
        v0 = torch.randint(0, BB.P, (1 << 23, 4), device=dev, dtype=torch.int32, generator=g)
        def fri4():
            gpu.fri_commit_phase(BB.id, _lib.HASH_KECCAK, v0.clone(), 1, 0, 3, 3, betas)
        t, nl = timed(fri4, k, 1)
        others["config4_fri_commit_phase_keccak_bb_2^23"] = {"ms": t, "launches": nl / k}
        del v0
        # config 5 hot path (prover.rs:215,319,394 minus the host-side AIR/quotient/opening work): KoalaBear, trace 2^20 x 1312,
        # blowup 2, Poseidon2-24 sponge + Poseidon2-16 compression, cap 3; quotient commit 2 x (2^20 x 4); FRI commit phase 2^21
        xt = torch.randint(0, KB.P, (1 << 20, 1312), device=dev, dtype=torch.int32, generator=g)
        t_trace, nl = timed(lambda: gpu.pcs_commit(KB.id, _lib.HASH_POSEIDON2_W24, xt, 1), 2, 1)
        tl, _ = timed(lambda: gpu.coset_lde_batch(KB.id, xt, 1, KB.generator), 2, 1)
        # pcs.open pre-FRI work on the resident LDE: two opening points (zeta, zeta*g) like uni-stark
        lde_t = gpu.coset_lde_batch(KB.id, xt, 1, KB.generator)
        del xt
        from plonky3_b200 import extension as X
        zs = [np.array([11, 22, 33, 44], dtype=np.uint32), np.array([55, 66, 77, 88], dtype=np.uint32)]
        al = np.array([5, 6, 7, 8], dtype=np.uint32)
        zinv0 = X.ef_inv(KB, zs[0])
        t_inv, _ = timed(lambda: gpu.open_inv_denoms(KB.id, 21, zs[0], zinv0), k, 1)
        invd, adj = gpu.open_inv_denoms(KB.id, 21, zs[0], zinv0)
        low = lde_t[: 1 << 20]
        t_col, _ = timed(lambda: gpu.columnwise_dot(KB.id, low, adj), k, 1)
        t_row, _ = timed(lambda: gpu.rowwise_dot(KB.id, lde_t, al), k, 1)
        rr = gpu.rowwise_dot(KB.id, lde_t, al); ro = torch.zeros((1 << 21, 4), dtype=torch.int32, device=dev)
        t_red, _ = timed(lambda: gpu.open_reduce(KB.id, ro, rr, invd, al, al), k, 1)
        col_bytes, row_bytes = (1 << 20) * 1312 * 4 + (1 << 20) * 16, (1 << 21) * 1312 * 4 + (1 << 21) * 16
        open_ms = 2 * (t_inv + t_col + t_red) + t_row
        others["config5_open_pre_fri_kb_2^21x1312"] = {
            "inv_denoms_ms": t_inv, "columnwise_dot_ms": t_col, "columnwise_dot_GBps": col_bytes / t_col / 1e6,
            "rowwise_dot_ms": t_row, "rowwise_dot_GBps": row_bytes / t_row / 1e6, "reduce_ms": t_red,
            "open_two_points_ms": open_ms, "hbm_peak_GBps": peak,
            "note": "barycentric evaluation at 2 points + alpha compression + quotient accumulation over the resident trace LDE; HBM-bound streaming reductions"}
        del lde_t, invd, adj, rr, ro, low
        xq = torch.randint(0, KB.P, (1 << 20, 4), device=dev, dtype=torch.int32, generator=g)
        def quot():
            a_ = gpu.coset_lde_batch(KB.id, xq, 1, KB.generator); b_ = gpu.coset_lde_batch(KB.id, xq, 1, KB.generator)
            gpu.merkle_commit(KB.id, _lib.HASH_POSEIDON2_W24, [a_, b_])
        t_quot, _ = timed(quot, k, 1)
        v1 = torch.randint(0, KB.P, (1 << 21, 4), device=dev, dtype=torch.int32, generator=g)
        kbetas = np.random.default_rng(3).integers(0, KB.P, size=(10, 4), dtype=np.uint32)
        def fri5():
            gpu.fri_commit_phase(KB.id, _lib.HASH_POSEIDON2_W24, v1.clone(), 1, 0, 3, 3, kbetas)
        t_fri, _ = timed(fri5, k, 1)
        others["config5_hot_path_kb_2^20x1312"] = {
            "commit_trace_ms": t_trace, "of_which_lde_ms": tl, "commit_quotient_ms": t_quot, "fri_commit_phase_ms": t_fri,
            "hot_path_total_ms": t_trace + t_quot + t_fri,
            "note": "device-resident LDE+Merkle+FRI of prove_prime_field_31 -f koala-bear -o poseidon-2-permutations -l 20; "
                    "AIR quotient evaluation and openings are host-side in the reference and out of scope (SURVEY 8f)"}
        del v1, xq
        line["others"] = others

    # ---- CPU baseline (rank 0, N=1)
    if world == 1 and rank == 0 and not args.no_cpu:
        try:
            v, cores, sample = cpu_lde_throughput()
            line["cpu_baseline"] = {"value": v, "unit": "Gelem/s", "cores": cores, "kind": "port", "sample": sample}
        except Exception as e:  # the oracle is test infrastructure; never let it break the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": "Gelem/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
