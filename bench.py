#!/usr/bin/env python
"""bench.py — headline benchmark of the erasure-coding + CRC engine (contract: task brief ④).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU; no data-path collective — chunks are independent,
   tiles are dealt round-robin, NCCL is used only for the barrier / max-over-ranks of the timing.)

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): ec(8,2) parity encode +
CRC32 of every 64 KiB data and parity block over 64 MiB chunks.  The 4096-chunk batch (256 GiB) does not
fit one B200, so a *step* is one pass of the hot path over one HBM-resident tile of --tile-chunks chunks
per GPU (default 512 = 32 GiB in, 8 GiB parity out); 8 steps on 1 GPU == the 4096-chunk batch.

One JSON line on stdout (rank 0).  `value` = chunk-data GiB/s with inputs resident in HBM; `e2e` = the
same metric through the host-buffer C-ABI call (pinned host memory, H2D + kernel + D2H inside the timed
region); `roofline` = algorithmic bytes / measured duration against the measured HBM peak;
`cpu_baseline` = the reference CPU implementation timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 65536
CHUNK = 64 << 20
GIB = float(1 << 30)
METRIC = "GiB/s encoded+CRC (ec(8,2), 64 MiB chunks)"


class stdout_to_stderr:
    """NCCL prints its version banner on fd 1 when the first communicator is created; the contract is ONE JSON line
    on stdout, so fd 1 is pointed at fd 2 while torch.distributed initialises."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--goal", default="ec(8,2)")
    ap.add_argument("--tile-chunks", type=int, default=512)
    ap.add_argument("--e2e-chunks", type=int, default=16)
    ap.add_argument("--cpu-chunks", type=int, default=0, help="chunks in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def algorithmic_bytes_per_chunk(k, m, chunk_len=CHUNK):
    """read S + write m*pb*B + write 4*(nb + m*pb)."""
    nb = (chunk_len + BLOCK - 1) // BLOCK
    pb = (nb + k - 1) // k
    return chunk_len + m * pb * BLOCK + 4 * (nb + m * pb)


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index):
        self.dev = device_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w": float(np.median(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(goal_text, n_chunks, threads, reps=1):
    """Times the reference CPU implementation (oracle/_ref when present, else the oracle port) on
    `n_chunks` synthetic 64 MiB chunks using `threads` host threads (chunk-level parallelism; ctypes
    releases the GIL).  Uses the reference CALL PATTERN (per stripe, per parity part rs.recover +
    mycrc32 per block: chunk_writer.cc:365-401, write_executor.cc:97).  Returns (GiB/s, kind, seconds)."""
    from concurrent.futures import ThreadPoolExecutor

    import ctypes as C

    from tests import _oracle as O
    import lizardfs_b200 as L  # only for goal parsing (host logic)

    g = L.SliceType(goal_text)
    oracle = O.load_oracle()
    ref = O.load_ref()
    lib, kind = (ref, "reference") if ref is not None else (oracle, "port")
    nb = CHUNK // BLOCK
    pb = (nb + g.k - 1) // g.k
    n_buf = min(n_chunks, max(threads, 1))
    chunks = [O.fill_chunk(oracle, CHUNK, 12345, c) for c in range(n_buf)]
    par = [np.zeros(g.m * pb * BLOCK, dtype=np.uint8) for _ in range(n_buf)]
    crc = [np.zeros(nb + g.m * pb, dtype=np.uint32) for _ in range(n_buf)]
    fn = lib.fn("encode_chunk")

    def one(i):
        b = i % n_buf
        rc = fn(g.kind, g.k, g.m, chunks[b].ctypes.data_as(C.c_void_p), C.c_size_t(CHUNK), par[b].ctypes.data_as(C.c_void_p),
                crc[b].ctypes.data_as(C.c_void_p))
        assert rc == 0
    one(0)  # warm-up (tables, page faults)
    best = None
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for _ in range(reps):
            t0 = time.perf_counter()
            list(ex.map(one, range(n_chunks)))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return n_chunks * CHUNK / GIB / best, kind, best


def cpu_reference_best_case(goal_text, n_chunks, threads):
    """SURVEY.md §8d(ii) "best-case reference kernel": ONE rs.encode over whole part-major parts
    (reed_solomon.h:134-155) + mycrc32 per block, parts prepared outside the timed region.  Only with oracle/_ref."""
    from concurrent.futures import ThreadPoolExecutor

    import ctypes as C

    from tests import _oracle as O
    import lizardfs_b200 as L

    ref = O.load_ref()
    if ref is None:
        return None
    g = L.SliceType(goal_text)
    oracle = O.load_oracle()
    nb = CHUNK // BLOCK
    pb = (nb + g.k - 1) // g.k
    n_buf = min(n_chunks, max(threads, 1))
    bufs = []
    for c in range(n_buf):
        parts, _ = O.split_parts(O.fill_chunk(oracle, CHUNK, 12345, c), g.k)
        parts = [np.ascontiguousarray(p) for p in parts]
        bufs.append((parts, O.ptr_array(parts), np.zeros(g.m * pb * BLOCK, dtype=np.uint8), np.zeros((g.k + g.m) * pb, dtype=np.uint32)))
    fn = ref.fn("encode_parts")

    def one(i):
        parts, ptrs, par, crc = bufs[i % n_buf]
        assert fn(g.kind, g.k, g.m, ptrs, C.c_uint32(pb), par.ctypes.data_as(C.c_void_p), crc.ctypes.data_as(C.c_void_p)) == 0
    one(0)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        t0 = time.perf_counter()
        list(ex.map(one, range(n_chunks)))
        dt = time.perf_counter() - t0
    return n_chunks * CHUNK / GIB / dt


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    n_chunks = args.cpu_chunks or max(cores, 8)
    # one "step" = a bounded sample of n_chunks chunks; warm-up steps are real runs as well
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(args.goal, min(n_chunks, cores), cores)
    vals, secs, kind = [], 0.0, "port"
    for _ in range(args.steps):
        v, kind, dt = cpu_reference_run(args.goal, n_chunks, cores)
        vals.append(v); secs += dt
    value = float(np.mean(vals))
    import lizardfs_b200 as L
    g = L.SliceType(args.goal)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GiB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{g} encode + per-64KiB-block CRC32, 64 MiB chunks, reference CPU call pattern "
                               f"(ChunkWriter::computeParityBlock per stripe/parity + mycrc32 per block)",
                   "chunks_per_step": n_chunks, "host_threads": cores},
        "cpu_baseline": {"value": value, "unit": "GiB/s", "cores": cores, "kind": kind,
                         "sample": f"{n_chunks} x 64 MiB chunks per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch

    import lizardfs_b200 as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # Host buffers of the e2e leg should live on the NUMA node next to this rank's GPU: bind the process to the GPU's
    # ideal CPUs before anything is allocated (restored before the CPU baseline, which must see every core).
    all_cpus = os.sched_getaffinity(0)
    numa_note = "not bound"
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        numa_note = f"bound to {len(os.sched_getaffinity(0))} CPUs local to GPU {phys}"
    except Exception as exc:  # no NVML / not permitted: keep the default placement
        numa_note = f"not bound ({type(exc).__name__})"
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
    dev = torch.device("cuda", local)
    eng = L.Engine(local)
    goal = L.SliceType(args.goal)
    k, m = goal.k, goal.m
    nb = CHUNK // BLOCK
    pb = (nb + k - 1) // k
    T = args.tile_chunks
    par_stride = m * pb * BLOCK
    crc_stride = nb + m * pb

    d_data = torch.empty(T * CHUNK, dtype=torch.uint8, device=dev)
    d_par = torch.empty(T * par_stride, dtype=torch.uint8, device=dev)
    d_crc = torch.empty(T * crc_stride, dtype=torch.int32, device=dev)
    # a dedicated non-default stream: handle 0 (torch's legacy default stream) would make the library
    # pick its own stream and the CUDA events below would not bracket the kernels
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    eng.fill_chunks_dev(d_data.data_ptr(), T, CHUNK, CHUNK, seed=12345, first_chunk=rank * T, stream=sptr)

    def step():
        eng.encode_chunks_dev(goal, T, CHUNK, d_data.data_ptr(), CHUNK, d_par.data_ptr(), par_stride, d_crc.data_ptr(), crc_stride, stream=sptr)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = eng.stats()["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    t_wall1 = time.time()
    launches = eng.stats()["kernel_launches"] - launches0
    ms_total = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    value = world * T * args.steps * CHUNK / GIB / (ms_total_max / 1e3)

    # spot check inside the bench: one chunk of this rank's tile against the CRC linearity identity
    crc_host = d_crc[:crc_stride].cpu().numpy().view(np.uint32)
    if goal.kind == 1 and not (m >= 5 or (m == 4 and k > 20)) and nb % k == 0:
        p0 = np.bitwise_xor.reduce(crc_host[:nb].reshape(pb, k), axis=1) ^ (np.uint32(0xD7978EEB) if k % 2 == 0 else np.uint32(0))
        if not os.environ.get("LZGPU_PROBE"):
            assert (p0 == crc_host[nb:nb + pb]).all(), "bench self-check failed: CRC(P) != xor of data CRCs"

    # ---- end-to-end through the host-buffer C-ABI call (pinned memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        E = args.e2e_chunks
        h_in = torch.empty(E * CHUNK, dtype=torch.uint8).pin_memory()
        h_par = torch.empty(E * par_stride, dtype=torch.uint8).pin_memory()
        h_crc = torch.empty(E * crc_stride, dtype=torch.int32).pin_memory()
        h_in.copy_(d_data[: E * CHUNK])
        np_in, np_par, np_crc = h_in.numpy(), h_par.numpy(), h_crc.numpy().view(np.uint32)
        import ctypes as C
        lib = eng.lib

        def e2e_step():
            rc = lib.lzgpu_encode_chunks(eng.h, C.byref(goal.c), E, CHUNK, np_in.ctypes.data_as(C.c_void_p), CHUNK,
                                         np_par.ctypes.data_as(C.c_void_p), par_stride, np_crc.ctypes.data_as(C.c_void_p), crc_stride)
            assert rc == 0, L._lib.last_error()
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        assert (np_crc[:crc_stride] == crc_host).all(), "e2e result differs from the resident run"
        e2e = {"value": world * E * args.steps * CHUNK / GIB / dt, "unit": "GiB/s", "h2d_bytes_per_step": E * CHUNK,
               "d2h_bytes_per_step": E * (par_stride + 4 * crc_stride), "chunks_per_step": E,
               "timing": "host wall clock around the synchronous C-ABI call, max over ranks", "host_placement": numa_note}

    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    launches_per_step = launches / max(1, args.steps)
    alg = algorithmic_bytes_per_chunk(k, m) * T
    kernel_ms = float(np.mean(step_ms))
    achieved = alg / (kernel_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_chunk") * T if tj.get("dram_bytes_per_chunk") else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                "kernel": "fused encode+CRC step" if launches_per_step <= 1.01 else f"{launches_per_step:.0f} kernels per step (generic route); duration = whole step",
                "launch_ms": kernel_ms}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)
        cores = host_threads()
        n = args.cpu_chunks or max(cores, 8)
        v, kind, secs = cpu_reference_run(args.goal, n, cores)
        best = cpu_reference_best_case(args.goal, n, cores)
        cpu_baseline = {"value": v, "unit": "GiB/s", "cores": cores, "kind": kind,
                        "sample": f"{n} x 64 MiB chunks ({secs:.1f} s), reference call pattern, {cores} threads",
                        "whole_part_kernel_value": best,
                        "whole_part_kernel_note": "one rs.encode over whole parts + mycrc32 per block (SURVEY 8d(ii)), parts pre-split outside the timed region"}

    line = {
        "metric": METRIC, "value": value, "unit": "GiB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{goal} encode + per-64KiB-block CRC32, 64 MiB chunks (BASELINE.json configs[2])",
                   "chunks_per_step_per_gpu": T, "resident_bytes_per_gpu": T * (CHUNK + par_stride + 4 * crc_stride),
                   "l2_policy": "inputs (32 GiB/step) far larger than the 126 MB L2; no flush needed",
                   "parallelism": f"static round-robin of chunk tiles over {world} GPU(s), no collective"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
