#!/usr/bin/env python
"""Benchmark of the Sequitr hot path on B200 (driver contract in the task statement).

Workload (BASELINE.json configs[2], the one `metric` is quoted on): UNet2D segmentation
(filters 16..256, concat bridge, 1 input channel, 2 classes) + connected-component
localisation on a synthetic 2048x2048 time-lapse, frames sharded across the ranks with no
collective on the data path.  One "step" = one batch of `--batch` frames through
UNet -> argmax mask -> label-and-localise -> centroid table.

  value     frames/s, whole job, inputs resident in HBM (device timed, CUDA events, max over ranks);
            the steps cycle through 4 batches of DISTINCT frames (32 frames, 537 MB)
  roofline  the UNet launches (21 tcgen05 conv / up-conv launches + the first conv, >99.9% of the FLOPs)
            timed with CUDA events around them INSIDE the timed steps: FLOPs / that time, against the
            sustained and the burst measured bf16 peaks; `frac_whole_step` divides by the whole step
  layers    per-layer CUDA-event pass run right after the timed loop (explanatory)
  parity    the benchmarked bf16 path against the fp32 exact mode (bit-identical to the oracle's fp32
            contract: tests/test_gpu_unet_bf16.py) on the same frames: mask mismatch fraction, the
            largest fp32 top-2 logit margin at a differing pixel, centroid rows changed; and the speed
            of that exact mode (`exact_mode_fps`), the price of bit-identical masks
  e2e       frames/s through the reference-facing call (UNet2D.segment_and_localise ->
            sq_segment_localise_raw_host) on camera-native uint16 HOST frames (what the reference's
            readers deliver, dataio/octopus.py:231-245) in, HOST tables out, all copies timed;
            `e2e_f32_host` is the same call fed float32 host frames
  stack2000 BASELINE configs[2] as written: a 2000-frame uint16 host stack, every rank streams its
            contiguous range through the host call; wall clock of the slowest shard; sha256 of the
            merged centroid tables (identical for every N) and a cross-rank check of a 64-frame prefix
  cpu_baseline  the oracle's CPU path (torch-CPU UNet restatement + SciPy label/centre-of-mass)
            timed on the host cores on a bounded sample (~10 s) of the same workload

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
SEED = 1234
LABEL_LAUNCHES = 7                         # kernels of one sq_label_centroids call (csrc/ccl.cu)
WORKLOAD = ("UNet2D seg + CCL localise, 2048x2048x1ch frames, filters 16-256, "
            "concat bridge, 2 classes (BASELINE configs[2])")


def dense_layer_traffic(filters=FILTERS, px=H * W, cin=1, classes_mask_bytes=1):
    """Compulsory HBM bytes per frame of every UNet launch, layer by layer, as (read, written) keyed by
    TF scope: each layer reads its bf16 input(s) once and writes its bf16 output once (DESIGN.md
    section 4); the first conv reads the fp32 frame; the last conv writes only the 1 B/px class mask."""
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
    # level-0 pairs that run as ONE launch on the quad layout (the intermediate stays in shared memory):
    # down0/conv1+conv2 reads the fp32 frame and writes the skip tensor + its pooled copy; up0/upscale+conv1
    # reads the level-1 tensor and the skip tensor and writes conv1's output
    f0 = filters[0]
    t['UNet/down0/conv1+conv2'] = (px * 4 * cin, px * 2 * f0 + (px / 4 * 2 * f0 if nl > 1 else 0))
    if nl > 1:
        t['UNet/up0/upscale+conv1'] = (px / 4 * 2 * filters[1] + px * 2 * f0, px * 2 * f0)
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
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_frames_per_s(max_frames, threads=None, budget_s=None):
    """The oracle CPU path on up to `max_frames` frames of the workload (stops early once `budget_s`
    seconds have been spent): returns (frames/s, threads, frames done)."""
    import torch
    from oracle import unet_oracle, centroid_oracle
    from sequitr_b200 import synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    weights = synth.blob_detector_weights(FILTERS, 1, 2, seed=1)
    done, spent = 0, 0.0
    while done < max_frames:
        n = min(2, max_frames - done)
        frames = synth.frames(n, H, W, 1, seed=SEED, first_frame=done)      # generation is not timed
        t0 = time.perf_counter()
        for i in range(n):
            out = unet_oracle.unet_forward(frames[i:i + 1], weights, FILTERS, 'concat')
            centroid_oracle.centroid_tables(out['mask'])
        spent += time.perf_counter() - t0
        done += n
        if budget_s is not None and spent >= budget_s:
            break
    return done / spent, threads, done


def run_reference(args, rank):
    """The reference arm: the CPU path of the oracle port on the host cores.  One process whatever
    --gpus says (a CPU baseline does not scale with the number of GPUs): rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.warmup > 0:
        cpu_reference_frames_per_s(1, cores)
    t0 = time.perf_counter()
    fps, threads, done = cpu_reference_frames_per_s(max(1, args.steps), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 / fps,
        "higher_is_better": True, "scaling": "weak", "scales_with_n": False, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": 1,
                   "note": "CPU baseline: ONE host process on all host threads whatever --gpus is; "
                           "per-N ratios against it are not scaling results"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frame(s) of 2048x2048: torch-CPU fp32 UNet restatement + "
                                   "SciPy label/center_of_mass (oracle/)" % done},
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
    ap.add_argument('--stack-frames', type=int, default=env_int('SQ_BENCH_STACK', 2000),
                    help='frames of the configs[2] time-lapse (0 = skip the stack2000 leg)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    args = ap.parse_args()

    rank, world, local = env_int('RANK', 0), env_int('WORLD_SIZE', 1), env_int('LOCAL_RANK', 0)
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    from sequitr_b200 import synth, shard
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    PB = 4                                                # distinct resident batches the steps cycle through
    CALL = 64                                             # frames per end-to-end host call
    t_wall0 = time.perf_counter()

    # ---- the time-lapse, BEFORE CUDA is initialised (forked renderers): this rank's contiguous range of
    #      the 2000-frame uint16 stack (frame t depends on (seed, t) only), at least PB*B + 2*CALL frames
    n_stack = max(args.stack_frames, 0)
    need = max(PB * B, 3 * CALL)
    total = max(n_stack, world * need)
    lo, hi = shard.frame_range(rank, world, total)
    gen_workers = max(1, (os.cpu_count() or 1) // max(1, env_int('LOCAL_WORLD_SIZE', world)))
    t0 = time.perf_counter()
    stack = synth.camera_stack(lo, hi, H, W, seed=SEED, workers=gen_workers)
    # the cross-rank prefix check: the LAST rank also renders global frames [0, CALL)
    prefix = None
    if n_stack and rank == world - 1:
        prefix = stack[:CALL] if world == 1 else synth.camera_stack(0, CALL, H, W, seed=SEED, workers=gen_workers)
    gen_s = time.perf_counter() - t0

    import torch
    import torch.distributed as dist
    import sequitr_b200
    from sequitr_b200 import ops, parity
    from sequitr_b200.networks import UNet2D

    torch.cuda.set_device(local)
    numa_cpus = shard.bind_to_gpu_numa(local) if world > 1 else None     # one worker per GPU, near its GPU
    host_group = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world,
                                device_id=torch.device('cuda', local))
        host_group = dist.new_group(backend='gloo')      # host-side table gathering only
    sequitr_b200.require_gpu(local)
    dev = torch.device('cuda', local)

    # page-lock the stack in place (cudaHostRegister) so the host call's copies are DMA at PCIe speed.  The OS can
    # refuse a multi-GB registration right after another process released one (the driver runs the arms back to
    # back): the registration is retried, and if it stays refused the stack is read at the staged-copy rate and the
    # frames of the e2e legs are copied into a pinned buffer -- the run never dies on it.  (One registration for the
    # whole stack: a copy may not span two separately registered ranges.)
    from sequitr_b200 import utils as sq_utils
    pinned_frac = 0.0 if os.environ.get('SQ_BENCH_NOPIN') else sq_utils.pin_in_place(stack, retries=3, wait_s=0.3)    # (the variable forces the fallback: tests)
    pinned_how = 'cudaHostRegister' if pinned_frac == 1.0 else 'cudaHostRegister refused (stack stays pageable)'
    pinned_head = None
    if pinned_frac < 1.0:
        try:                                             # 1st fallback: a page-locked allocation of the whole stack
            if os.environ.get('SQ_BENCH_NOPIN') == '2':
                raise MemoryError('forced')
            full = sq_utils.pinned_array(stack.shape, np.uint16)
            full[...] = stack
            stack = full
            pinned_frac = 1.0
            pinned_how += '; stack copied into a cudaHostAlloc buffer instead'
        except Exception:
            try:                                         # 2nd: only the frames of the e2e legs
                head = sq_utils.pinned_array((min(len(stack), 6 * CALL),) + stack.shape[1:], np.uint16)
                head[...] = stack[:len(head)]
                pinned_head = head
                pinned_how += '; first %d frames copied into a pinned buffer for the e2e legs' % len(head)
            except Exception:
                pinned_head = None

    def frames_for_calls(s, n):
        """Host frames [s, s+n) of this rank's stack: from the pinned head copy when only that is page-locked."""
        if pinned_head is not None and s + n <= len(pinned_head):
            return pinned_head[s:s + n]
        return stack[s:s + n]

    net = UNet2D({'filters': FILTERS, 'shape': (H, W), 'bridge': 'concat', 'num_inputs': 1,
                  'num_outputs': 2, 'compute': 'bf16', 'device': local})
    weights = synth.blob_detector_weights(FILTERS, 1, 2, seed=1)
    net.load_weights(weights)
    max_rows = 2048
    label_ws = ops.Workspace(dev)

    # ---- resident inputs: PB batches of distinct frames, widened + normalised (ImageNorm) once, fp32
    pool_raw = torch.from_numpy(stack[:PB * B].astype(np.int32)).to(dev).to(torch.float32).unsqueeze(-1)
    dev_pool = ops.image_norm(pool_raw.contiguous())
    del pool_raw

    def device_step(i, frame0, ev=None):
        xb = dev_pool[(i % PB) * B:(i % PB + 1) * B]
        if ev is not None:
            ev[0].record()
        mask = net.predict(xb, want=('mask',))['mask']
        if ev is not None:
            ev[1].record()
        table, counts = ops.label_centroids(mask, max_rows=max_rows, frame0=frame0, workspace=label_ws)
        return table, counts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: inputs resident in HBM
    for i in range(max(Wm, PB)):
        table, counts = device_step(i, lo + i * B)
    barrier()
    n_obj = int(counts.sum().item())
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        table, counts = device_step(i, lo + i * B, evs[i])
    e1.record()
    barrier()
    clocks = sampler.result()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * K / (ms_total * 1e-3)
    unet_launches = net.launches()
    launches_per_step = unet_launches + LABEL_LAUNCHES
    unet_ms = sum(a.elapsed_time(b) for a, b in evs)     # this rank's UNet launches inside the timed steps

    # ---- per-layer pass, immediately after the timed loop (same thermal / power state)
    layer_ms, layer_fl, reps = {}, {}, 3
    if rank == 0:
        for _ in range(reps):
            for name, lms, fl in net.profile(dev_pool[:B]):
                layer_ms[name] = layer_ms.get(name, 0.0) + lms / reps
                layer_fl[name] = fl

    # ---- roofline, from the timed region itself
    roofline = roofline_hbm = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        sustained, burst = peaks.get('bf16_tflops_sustained'), peaks.get('bf16_tflops')
        which = 'MEASURED_PEAKS.json'
        if not sustained or not burst:
            sustained, burst, which = 1400.0, 1700.0, 'fallback (B200_PROFILING.md)'
        hbm_peak = peaks.get('hbm_gbs') or 6650.0
        step_ms_rank0 = e0.elapsed_time(e1) / K
        assert unet_ms / K <= step_ms_rank0 * 1.0005, (unet_ms / K, step_ms_rank0)
        flops_step = FLOP_PER_FRAME * B
        achieved = flops_step * K / (unet_ms * 1e-3) / 1e12
        capped = 'sw_power_cap' in clocks.get('reasons', [])
        at_max = clocks.get('sm_mhz') and clocks.get('sm_max_mhz') and clocks['sm_mhz'] >= 0.97 * clocks['sm_max_mhz']
        regime = 'burst (clocks at max, no power cap seen)' if (at_max and not capped) else \
                 'power-capped / sustained' if capped else 'below max clocks'
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'conv_traffic.json')))['avg_dram_bytes_per_launch']
        except Exception:
            traffic = None
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
            "frac": achieved / sustained, "traffic": traffic,
            "frac_vs_sustained": achieved / sustained, "frac_vs_burst": achieved / burst,
            "peak_burst": burst, "peak_source": which, "regime": regime,
            "frac_whole_step": flops_step / (step_ms_rank0 * 1e-3) / 1e12 / sustained,
            "frac_whole_step_vs_burst": flops_step / (step_ms_rank0 * 1e-3) / 1e12 / burst,
            "kernel": "the UNet launches of a step (tcgen05 family conv_tc / conv_xc / conv_qd / conv_qf / conv_qu: "
                      "%d launches), CUDA events around them inside the timed steps" % unet_launches,
            "unet_ms_per_step": unet_ms / K, "step_ms": step_ms_rank0,
            "avg_launch_ms": unet_ms / K / max(unet_launches, 1),
            "flops_per_step": flops_step,
        }
        # the same launches against HBM, layer by layer (explanatory): compulsory bytes of every layer
        tr = dense_layer_traffic()
        # the launches this run actually made (fused level-0 pairs appear under "a+b" names)
        tr = {k: v for k, v in tr.items() if k in layer_ms} or tr
        alg_bytes = sum(rd + wr for rd, wr in tr.values()) * B
        floor_ms = meas_ms = 0.0
        for name, (rd, wr) in tr.items():
            if name in layer_ms:
                floor_ms += 1e3 * max((rd + wr) * B / (hbm_peak * 1e9), layer_fl.get(name, 0.0) / (sustained * 1e12))
                meas_ms += layer_ms[name]
        roofline_hbm = {
            "bound": "hbm", "achieved": alg_bytes / (unet_ms / K * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": alg_bytes / (unet_ms / K * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
            "algorithmic_bytes_per_step": alg_bytes,
            "kernel": "same launches, compulsory activation bytes of the schedule that ran (fused level-0 pairs "
                      "keep their intermediate on chip)",
            "algorithmic_bytes_per_frame": alg_bytes / B,
            "layer_floor": {"floor_ms_per_step": floor_ms, "measured_ms_per_step": meas_ms,
                            "frac": floor_ms / meas_ms if meas_ms else None,
                            "model": "sum over the UNet launches of max(bytes / HBM copy peak, FLOPs / sustained "
                                     "bf16 peak), per-layer pass run right after the timed loop"},
        }

    # ---- parity of the benchmarked path against the fp32 exact mode, and the price of exactness
    par = None
    if rank == 0 and not args.no_parity:
        NP = 2
        xs = dev_pool[:NP]
        a = net.predict(xs, want=('mask',))
        net32 = UNet2D({'filters': FILTERS, 'shape': (H, W), 'bridge': 'concat', 'num_inputs': 1,
                        'num_outputs': 2, 'compute': 'fp32', 'device': local})
        net32.load_weights(weights)
        b = net32.predict(xs, want=('logits', 'mask'))
        torch.cuda.synchronize()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        m32 = net32.predict(xs, want=('mask',))['mask']
        ops.label_centroids(m32, max_rows=max_rows, frame0=0, workspace=label_ws)
        x1.record()
        torch.cuda.synchronize()
        exact_fps = NP / (x0.elapsed_time(x1) * 1e-3)
        mp = parity.mask_parity(a['mask'].cpu().numpy(), b['mask'].cpu().numpy(), b['logits'].cpu().numpy())

        def tables_of(mask):
            t, c = ops.label_centroids(mask, max_rows=max_rows, frame0=0, workspace=label_ws)
            t, c = t.cpu().numpy(), c.cpu().numpy()
            return [t[i, :c[i]] for i in range(len(c))]
        cd = parity.centroid_set_diff(tables_of(a['mask']), tables_of(b['mask']), tol_px=0.5)
        par = {"against": "fp32 exact mode (bit-identical to the oracle's fp32 contract), %d frames of the "
                          "resident pool" % NP,
               "mask_mismatch_frac": mp['mismatch_frac'], "mask_mismatch_pixels": mp['mismatch'],
               "max_margin_of_mismatch": mp['max_margin_of_mismatch'], "logit_range": mp['logit_range'],
               "centroid_rows": cd['rows'], "centroid_rows_ref": cd['ref_rows'],
               "centroid_rows_changed": cd['rows_changed'], "centroid_rows_identical": cd['identical'],
               "centroid_rows_unmatched": cd['unmatched'] + cd['ref_unmatched'],
               "centroid_max_shift_px": cd['max_shift_px'],
               "exact_mode_fps": exact_fps,
               "note": "bit-identical masks are delivered by compute='fp32' at exact_mode_fps; the bf16 path "
                       "flips only near-ties of the fp32 logits (margin rule asserted in tests/)"}
        del net32, a, b, m32
        torch.cuda.empty_cache()

    # ---- the training step of BASELINE configs[4] at tr_augment's default crop (4 crops of 512x512), aux key
    train_res = None
    if rank == 0 and world == 1 and not args.no_parity:
        from sequitr_b200.networks.unet import ModeKeys
        tnet = UNet2D({'filters': FILTERS, 'shape': (512, 512), 'bridge': 'concat', 'num_inputs': 1,
                       'num_outputs': 2, 'compute': 'fp32', 'dropout': 0.4, 'device': local}, mode=ModeKeys.TRAIN)
        tnet.load_weights(synth.unet_weights(FILTERS, 1, 2, ndim=2, bridge='concat', seed=1))
        trainer = tnet.trainer(learning_rate=1e-3, seed=1)
        timg = dev_pool[:1, :1024, :1024].reshape(1, 2, 512, 2, 512, 1).permute(0, 1, 3, 2, 4, 5).reshape(4, 512, 512, 1)
        timg = timg.contiguous()
        tlab = (timg[..., 0] > timg.mean()).to(torch.uint8).contiguous()
        twgt = ops.weightmap_edt(tlab, 10., 5.).float()
        losses = [trainer.step(timg, tlab, twgt) for _ in range(2)]          # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        TS = 5
        for _ in range(TS):
            losses.append(trainer.step(timg, tlab, twgt))                    # the loss read-back synchronises
        train_ms = (time.perf_counter() - t0) / TS * 1e3
        train_res = {"ms_per_step": train_ms, "crops_per_s": 4 / train_ms * 1e3,
                     "batch": "4 crops of 512x512x1 (tr_augment's default shape), dropout 0.4, Adam",
                     "arithmetic": "fp32 CUDA cores (tiled conv / dgrad, split-pixel wgrad); not on tensor cores",
                     "loss_first": losses[0], "loss_last": losses[-1]}
        trainer.close()
        del tnet, trainer
        torch.cuda.empty_cache()

    # ---- e2e: the reference-facing host call on camera-native uint16 frames, host tables out.
    #      One call = CALL frames = CALL/B steps (a Sequitr job hands a stack to the network, not 8 frames
    #      at a time); inside, frames stream in chunks of 1, 1, 2, 4, 8, 8, ... (H2D of a chunk under the
    #      UNet of the previous one); widening + ImageNorm run on the device.  Every call reads DIFFERENT frames.
    nloc = hi - lo
    net.segment_and_localise(frames_for_calls(0, CALL), frame0=lo, max_rows=max_rows, normalise=True)
    net.segment_and_localise(frames_for_calls(0, CALL), frame0=lo, max_rows=max_rows, normalise=True)
    ncalls = max(2, min(4, nloc // CALL - 1))
    barrier()
    t0 = time.perf_counter()
    for c in range(ncalls):
        s = ((c + 1) * CALL) % max(nloc - CALL + 1, 1)
        net.segment_and_localise(frames_for_calls(s, CALL), frame0=lo + s, max_rows=max_rows, normalise=True)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * ncalls * CALL / e2e_s
    d2h = B * max_rows * 5 * 4 + B * 4

    # ---- secondary: the same call fed float32 host frames (4 B/px over PCIe, no device-side ImageNorm)
    try:
        f32 = torch.empty((CALL, H, W, 1), dtype=torch.float32).pin_memory()
    except Exception:                                    # the OS refused the page-locked allocation: pageable
        torch.cuda.synchronize()
        f32 = torch.empty((CALL, H, W, 1), dtype=torch.float32)
    f32.numpy()[..., 0] = (stack[:CALL].astype(np.float32) - 3000.0) / 400.0
    f32_np = f32.numpy()
    net.segment_and_localise(f32_np, frame0=lo, max_rows=max_rows)
    barrier()
    t0 = time.perf_counter()
    for c in range(2):
        net.segment_and_localise(f32_np, frame0=lo, max_rows=max_rows)
    torch.cuda.synchronize()
    e2e_f32 = world * 2 * CALL / max_over_ranks(time.perf_counter() - t0)
    del f32, f32_np

    # ---- stack2000: configs[2] as written
    stack_res = None
    if n_stack:
        per_call = 250
        # untimed warm-up call of the same size: the first call of a size grows the handle's device arena
        # (one cudaMalloc of several GB), which a time-lapse job pays once, not per 250 frames
        net.segment_and_localise(stack[:min(per_call, nloc)], frame0=lo, max_rows=max_rows, normalise=True)
        barrier()
        t0 = time.perf_counter()
        tables = shard.segment_stack(net, stack[:nloc], frame0=lo, frames_per_call=per_call,
                                     max_rows=max_rows, normalise=True)
        torch.cuda.synchronize()
        wall = max_over_ranks(time.perf_counter() - t0)
        # cross-rank check: the last rank recomputes global frames [0, CALL) in ONE call (other GPU, other
        # chunking, other neighbours in the batch) and rank 0 compares them with its own tables
        pre = None
        if rank == world - 1:
            pre = net.segment_and_localise(prefix, frame0=0, max_rows=max_rows, normalise=True)
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, (tables, pre), group=host_group)
            all_tables = shard.merge_tables([g[0] for g in gathered])
            pre = gathered[-1][1]
        else:
            all_tables = tables
        if rank == 0:
            assert len(all_tables) == total
            prefix_equal = all(np.array_equal(a, b) for a, b in zip(pre, all_tables[:CALL]))
            frames_idx_ok = all((len(t) == 0 or (t[:, 0] == i).all()) for i, t in enumerate(all_tables))
            stack_res = {"frames": total, "wall_s": wall, "value": total / wall, "unit": "frames/s",
                         "frames_per_rank": nloc, "frames_per_call": per_call, "input": "uint16 host stack",
                         "tables_sha256": shard.tables_digest(all_tables),
                         "objects": int(sum(len(t) for t in all_tables)),
                         "prefix64_equal_across_ranks": bool(prefix_equal),
                         "frame_column_global": bool(frames_idx_ok),
                         "timing": "wall clock of the slowest shard (max over ranks), host frames in, host "
                                   "tables out, after one untimed call of frames_per_call frames (arena growth); "
                                   "table merge is host-side and not timed"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps_cpu, threads, done = cpu_reference_frames_per_s(32, budget_s=10.0)
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "%d frames of 2048x2048 (>= 10 s of CPU work): torch-CPU fp32 UNet restatement + SciPy "
                         "label/center_of_mass (oracle/), %d host threads" % (done, threads)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": max(Wm, PB), "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "frames_per_step_per_gpu": B, "objects_last_step": n_obj,
                       "distinct_resident_frames": PB * B,
                       "l2": "inputs (134 MB/step, 4 distinct batches) and activations (>10 GB/step) exceed the 126 MB L2",
                       "sharding": "contiguous frame range per rank, no collective",
                       "worker_cpus": len(numa_cpus) if numa_cpus else None,
                       "stack_generation_s": gen_s, "stack_pinned": pinned_how},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": B * H * W * 2,
                    "d2h_bytes_per_step": d2h, "steps": ncalls * CALL // B, "steps_per_call": CALL // B,
                    "input": "uint16 camera frames in page-locked host memory (dataio/octopus.py:231-245), "
                             "widened + ImageNorm on the device; host centroid tables out"},
            "e2e_f32_host": {"value": e2e_f32, "unit": "frames/s", "h2d_bytes_per_step": B * H * W * 4,
                             "d2h_bytes_per_step": d2h, "input": "float32 host frames"},
            "stack2000": stack_res,
            "gpu_launches": launches_per_step * K,
            "roofline": roofline,
            "roofline_hbm": roofline_hbm,
            "layers_ms_per_step": {k: round(v, 4) for k, v in layer_ms.items()},
            "parity": par,
            "train_step": train_res,
            "cpu_baseline": cpu,
            "wall_s": time.perf_counter() - t_wall0,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
