#!/usr/bin/env python
"""Benchmark of the Sequitr hot path on B200 (driver contract in the task statement).

Workload (BASELINE.json configs[2], the one `metric` is quoted on): UNet2D segmentation
(filters 16..256, concat bridge, 1 input channel, 2 classes) + connected-component
localisation on a synthetic 2048x2048 time-lapse, frames sharded across the ranks with no
collective on the data path.  One "step" = one batch of `--batch` frames through
UNet -> argmax mask -> label-and-localise -> centroid table.

  value : frames/s, whole job, inputs resident in HBM (device timed, CUDA events, max over ranks)
  e2e   : frames/s through the reference-facing call (UNet2D.segment_and_localise ->
          sq_segment_localise_host) with pinned HOST frames in and HOST tables out, copies timed
  roofline : the tcgen05 conv/up-conv kernel family (the only tensor-core kernels, ~99% of the
          FLOPs): algorithmic FLOPs / CUDA-event time of those launches inside a step
  cpu_baseline : the oracle's CPU path (torch-CPU UNet restatement + SciPy label/centre-of-mass)
          timed on the host cores on a bounded sample of the same workload

`--impl reference` times that CPU path alone (the reference itself is Python-2/TF-1 and
cannot run) and prints the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FILTERS = (16, 32, 64, 128, 256)
H = W = 2048
FLOP_PER_FRAME = 92000.0 * H * W


def dense_layer_bytes(filters=FILTERS, px=H * W, classes_mask_bytes=1):
    """Algorithmic (compulsory) HBM bytes of the tcgen05 layers per frame, layer by layer: every layer
    reads its bf16 input(s) once and writes its bf16 output once; the last conv
    writes only the 1 B/px class mask (fused head).  The first conv (CUDA cores) is not included."""
    nl, total = len(filters), 0.0
    for l, f in enumerate(filters):
        p = px / 4 ** l
        if l > 0:
            total += p * 2 * (filters[l - 1] + f)                       # down{l}/conv1
        total += p * 2 * (f + f) + (p / 4 * 2 * f if l < nl - 1 else 0)    # conv2 (+ pooled copy)
    for l in range(nl - 2, -1, -1):
        p, f = px / 4 ** l, filters[l]
        total += p / 4 * 2 * filters[l + 1] + p * 2 * f                # upscale
        total += p * 2 * (2 * f + f)                                   # conv1 on concat(up, skip)
        total += p * 2 * f + (p * 2 * f if l > 0 else p * classes_mask_bytes)
    return total


def dense_layer_traffic(filters=FILTERS, px=H * W, cin=1, classes_mask_bytes=1):
    """The same compulsory bytes split per layer into (read, written), keyed by TF scope (the first
    conv included): the per-layer floors of the bench line are built from these."""
    nl, t = len(filters), {}
    for l, f in enumerate(filters):
        p = px / 4 ** l
        t['UNet/down%d/conv1' % l] = (p * (4 * cin if l == 0 else 2 * filters[l - 1]), p * 2 * f)
        t['UNet/down%d/conv2' % l] = (p * 2 * f, p * 2 * f + (p / 4 * 2 * f if l < nl - 1 else 0))
    for l in range(nl - 2, -1, -1):
        p, f = px / 4 ** l, filters[l]
        t['UNet/up%d/upscale' % l] = (p / 4 * 2 * filters[l + 1], p * 2 * f)
        t['UNet/up%d/conv1' % l] = (p * 2 * 2 * f, p * 2 * f)
        t['UNet/up%d/conv2' % l] = (p * 2 * f, p * 2 * f if l > 0 else p * classes_mask_bytes)
    return t


METRIC = "frames/sec 2048^2 UNet2D seg+localize"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons DURING the timed region (NVML)."""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwPowerBrakeSlowdown', 0x80): 'hw_power_brake',
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def cpu_reference_frames_per_s(n_frames, threads=None):
    """The oracle CPU path on `n_frames` frames of the workload: returns (frames/s, threads)."""
    import torch
    from oracle import unet_oracle, centroid_oracle
    from sequitr_b200 import synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    weights = synth.blob_detector_weights(FILTERS, 1, 2, seed=1)
    frames = synth.frames(n_frames, H, W, 1, seed=1234)
    t0 = time.perf_counter()
    for i in range(n_frames):
        out = unet_oracle.unet_forward(frames[i:i + 1], weights, FILTERS, 'concat')
        centroid_oracle.centroid_tables(out['mask'])
    dt = time.perf_counter() - t0
    return n_frames / dt, threads


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.warmup > 0:
        cpu_reference_frames_per_s(1, cores)
    t0 = time.perf_counter()
    fps, threads = cpu_reference_frames_per_s(max(1, args.steps), cores)
    ms_per_step = 1e3 / fps
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "UNet2D seg + CCL localise, 2048x2048x1ch frames, filters 16-256, "
                               "concat bridge, 2 classes (BASELINE configs[2]); one step = 1 frame"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frame(s) of 2048x2048: torch-CPU fp32 UNet restatement + "
                                   "SciPy label/center_of_mass (oracle/)" % max(1, args.steps)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=8, help='frames per step per GPU')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()

    rank, world, local = env_int('RANK', 0), env_int('WORLD_SIZE', 1), env_int('LOCAL_RANK', 0)
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import sequitr_b200
    from sequitr_b200 import synth, ops, shard
    from sequitr_b200.networks import UNet2D

    torch.cuda.set_device(local)
    numa_cpus = shard.bind_to_gpu_numa(local) if world > 1 else None     # one worker per GPU, near its GPU
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world,
                                device_id=torch.device('cuda', local))
    sequitr_b200.require_gpu(local)
    dev = torch.device('cuda', local)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)

    # ---- synthetic time-lapse: every rank owns a contiguous frame range of the whole job
    total_frames = world * B * (K + Wm)
    lo, hi = shard.frame_range(rank, world, total_frames)
    pool_n = B                                            # distinct frames kept resident per rank
    host_pool = torch.empty((pool_n, H, W, 1), dtype=torch.float32).pin_memory()
    host_pool.numpy()[...] = synth.frames(pool_n, H, W, 1, seed=1234, first_frame=lo % 1000)
    dev_pool = host_pool.to(dev, non_blocking=False)      # 134 MB of inputs (> 126 MB L2)

    net = UNet2D({'filters': FILTERS, 'shape': (H, W), 'bridge': 'concat', 'num_inputs': 1,
                  'num_outputs': 2, 'compute': 'bf16'})
    net.load_weights(synth.blob_detector_weights(FILTERS, 1, 2, seed=1))
    max_rows = 2048
    label_ws = ops.Workspace(dev)

    def device_step(frame0):
        mask = net.predict(dev_pool, want=('mask',))['mask']
        table, counts = ops.label_centroids(mask, max_rows=max_rows, frame0=frame0, workspace=label_ws)
        return table, counts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    for i in range(Wm):
        table, counts = device_step(lo + i * B)
    barrier()
    n_obj = int(counts.sum().item())
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        table, counts = device_step(lo + (Wm + i) * B)
    e1.record()
    barrier()
    clocks = sampler.result()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * K / (ms_total * 1e-3)
    launches_per_step = net.launches() + 10                # + the 10 label-and-localise kernels

    # ---- e2e: the reference-facing host call, pinned host frames in, host tables out.
    #      One call covers CALL_STEPS steps = 64 frames (a Sequitr job hands a whole stack -- the
    #      north star's is 2000 frames -- to the network, not 8 frames at a time); inside, frames stream
    #      through in chunks of 1, 1, 2, 4, 8, 8, ... so the H2D copy of a chunk overlaps the UNet of the
    #      previous one.  Every step's frames are copied
    #      from pinned host memory and every step's centroid table is read back.
    CALL_STEPS = int(os.environ.get('SQ_BENCH_CALL_STEPS', 8))
    big = torch.empty((CALL_STEPS * B, H, W, 1), dtype=torch.float32).pin_memory()
    for j in range(CALL_STEPS):
        big[j * B:(j + 1) * B] = host_pool
    frames_np = big.numpy()
    for _ in range(2):
        net.segment_and_localise(frames_np, frame0=lo, max_rows=max_rows)
    barrier()
    Kc = max(2, min(K, 12) // CALL_STEPS)
    Ke = Kc * CALL_STEPS
    t0 = time.perf_counter()
    for i in range(Kc):
        tables = net.segment_and_localise(frames_np, frame0=lo + i * CALL_STEPS * B, max_rows=max_rows)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * B * Ke / float(dt.item())
    h2d = B * H * W * 4
    d2h = B * max_rows * 5 * 4 + B * 4

    # ---- same call on RAW uint16 camera frames (what dataio.OctopusData.frames_raw delivers): 2 bytes
    #      per pixel cross PCIe, widening + ImageNorm run on the device (informational, not the headline)
    raw = torch.empty((CALL_STEPS * B, H, W), dtype=torch.uint16).pin_memory()
    raw.numpy()[...] = np.clip(frames_np[..., 0] * 400.0 + 3000.0, 0, 65535).astype(np.uint16)
    raw_np = raw.numpy()
    for _ in range(2):
        net.segment_and_localise(raw_np, frame0=lo, max_rows=max_rows, normalise=True)
    barrier()
    t0 = time.perf_counter()
    for i in range(Kc):
        net.segment_and_localise(raw_np, frame0=lo + i * CALL_STEPS * B, max_rows=max_rows, normalise=True)
    torch.cuda.synchronize()
    dtr = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dtr, op=dist.ReduceOp.MAX)
    e2e_raw = world * B * Ke / float(dtr.item())

    # ---- roofline of the dominant (tensor-core) kernel family: per-layer CUDA events on the
    #      stream the kernels run on, same batch as the timed step
    roofline = roofline_hbm = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak, which = peaks.get('bf16_tflops_sustained'), 'measured (sustained)'
        if not peak:
            peak, which = 1400.0, 'fallback'
        dense_ms, dense_fl, n_dense = 0.0, 0.0, 0
        for _ in range(3):
            rows = net.profile(dev_pool)
        reps = 3
        layer_ms = {}
        for _ in range(reps):
            for name, lms, fl in net.profile(dev_pool):
                layer_ms[name] = layer_ms.get(name, 0.0) + lms / reps
                if fl > 0 and name not in ('UNet/down0/conv1', 'UNet/to_image'):
                    dense_ms += lms
                    dense_fl += fl
                    n_dense += 1
        achieved = dense_fl / (dense_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": None,
                    "kernel": "conv_tc_kernel / conv_xc_kernel family (tcgen05 3x3 conv + up-conv, %d launches "
                              "per step)" % (n_dense // reps),
                    "peak_source": which,
                    "avg_launch_ms": dense_ms / n_dense,
                    "dense_ms_per_step": dense_ms / reps,
                    "flops_per_step": dense_fl / reps}
        # The same launches against HBM: layer by layer this net is bandwidth-bound at levels 0-1
        # (64-128 B moved per pixel for 4.6-37 kFLOP), so the byte roofline is reported next to the
        # tensor one.  traffic = ncu dram bytes per launch of the same family (profiles/).
        hbm_peak = peaks.get('hbm_gbs') or 6650.0
        alg_bytes = dense_layer_bytes() * B
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r1_conv_traffic.json')))['avg_dram_bytes_per_launch']
        except Exception:
            traffic = None
        roofline["traffic"] = traffic
        roofline_hbm = {"bound": "hbm", "achieved": alg_bytes / (dense_ms / reps * 1e-3) / 1e9, "peak": hbm_peak,
                        "unit": "GB/s", "frac": alg_bytes / (dense_ms / reps * 1e-3) / 1e9 / hbm_peak,
                        "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes / (n_dense // reps),
                        "kernel": "same 21 launches, compulsory bf16 activation bytes (each layer reads its "
                                  "inputs once, writes its output once)"}

        # Per-layer floor: max(compulsory bytes / HBM copy peak, FLOPs / sustained tensor peak), summed over
        # every UNet launch of a step (first conv included) against the time those launches took: how far
        # the layer-by-layer design is from ITS OWN limits (explanatory; `roofline` above is the headline).
        floor_ms = meas_ms = 0.0
        flops_l = {name: fl for name, _, fl in rows}
        for name, (rd, wr) in dense_layer_traffic().items():
            if name in layer_ms:
                floor_ms += 1e3 * max((rd + wr) * B / (hbm_peak * 1e9), flops_l.get(name, 0.0) / (peak * 1e12))
                meas_ms += layer_ms[name]
        roofline_hbm["layer_floor"] = {
            "floor_ms_per_step": floor_ms, "measured_ms_per_step": meas_ms,
            "frac": floor_ms / meas_ms if meas_ms else None,
            "model": "sum over the UNet launches of max(bytes / HBM copy peak, FLOPs / sustained bf16 peak)"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps_cpu, threads = cpu_reference_frames_per_s(2)
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "2 frames of 2048x2048: torch-CPU fp32 UNet restatement + SciPy "
                         "label/center_of_mass (oracle/), %d host threads" % threads}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "UNet2D seg + CCL localise, 2048x2048x1ch frames, filters 16-256, "
                                   "concat bridge, 2 classes (BASELINE configs[2])",
                       "frames_per_step_per_gpu": B, "objects_per_step": n_obj,
                       "l2": "inputs (134 MB/step) and activations (>10 GB/step) exceed the 126 MB L2",
                       "sharding": "contiguous frame range per rank, no collective",
                       "worker_cpus": len(numa_cpus) if numa_cpus else None},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": Ke, "steps_per_call": CALL_STEPS},
            "e2e_raw_u16": {"value": e2e_raw, "unit": "frames/s", "h2d_bytes_per_step": B * H * W * 2,
                            "d2h_bytes_per_step": d2h,
                            "note": "uint16 host frames, widened + ImageNorm on the device"},
            "gpu_launches": launches_per_step * K,
            "roofline": roofline,
            "roofline_hbm": roofline_hbm if roofline else None,
            "cpu_baseline": cpu,
            "tensor_frac_of_frame_flops": value / world * FLOP_PER_FRAME / 1e12 / (roofline or {}).get("peak", 1.0)
            if roofline else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
