#!/usr/bin/env python
"""bench.py -- decode Mpixels/s of the FUIF hot path on B200 (BASELINE.json's metric).

A step = one full decode of the workload: fuif_decode (MANIAC range decoder + context model on the GPU) followed by
Image::undo_transforms (inverse Squeeze / DCT / colour transforms on the GPU), bit-exact with the reference.

  value  : whole-job Mpx/s with the compressed file already resident in HBM and the pixels left in HBM
  e2e    : the same through the host-buffer C-ABI call fb_decode_to_pixels (pinned host bytes in, pinned host pixels
           out, H2D / D2H inside the timed region)
  roofline      : the dominant kernel of the inverse transform chain (the last launch: final horizontal unsqueeze step of the
                  chroma planes + inverse YCoCg + clamp), timed with CUDA events on the library's stream inside every timed step:
                  its algorithmic bytes (every input coefficient read once + every output sample written once) / time;
                  the whole chain (4*W*H*C bytes / undo_transforms time) is reported next to it
  cpu_baseline  : the reference's own CPU decoder (oracle/_ref/ref_driver, else the C port) on this box's host cores
  --impl reference : times the reference CPU decoder on the same workload and prints the same JSON line.

Launch: python bench.py --gpus N --steps K --warmup W     (N>1 via torch.distributed.run, one rank per GPU)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-index", action="store_true", help="decode without the group-offset sidecar (one stream per image)")
    ap.add_argument("--no-index-steps", type=int, default=1,
                    help="besides the main (indexed) measurement, time this many steps WITHOUT the sidecar and report value_no_index / e2e_no_index (0 = skip)")
    ap.add_argument("--distinct-images", action="store_true", help="weak scaling with a different image (seed) on every rank instead of replicas of the same one")
    ap.add_argument("--strong", action="store_true", help="multi-image workloads (cfg4): split the fixed batch over the ranks (strong scaling)")
    ap.add_argument("--host-entropy-steps", type=int, default=1,
                    help="also time this many e2e steps with the host-threads entropy backend and report e2e_host_entropy (0 = skip)")
    ap.add_argument("--host-threads", type=int, default=0, help="threads of the host entropy backend per GPU (0 = 4 x host cores / GPUs)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU decoder on the same workload, one process per image."""
    import bench_workloads as wl
    if rank != 0:
        return
    per = wl.WORKLOADS[args.workload][5]
    n_img = per if args.strong else max(1, args.gpus) * per
    seeds = range(n_img) if (args.distinct_images or args.strong) else [u % per for u in range(n_img)]     # replicas: the same image(s) for every GPU
    imgs = wl.prepare_images(args.workload, sorted(set(seeds)), want_index=False)
    by_seed = {u: im for u, im in zip(sorted(set(seeds)), imgs)}
    imgs = [by_seed[u] for u in seeds]
    mpix = sum(im["w"] * im["h"] for im in imgs) / 1e6
    use_ref = wl.have_ref_driver()

    def one_step():
        t0 = time.perf_counter()
        if use_ref:
            procs = [subprocess.Popen([wl.REF_DRIVER, "time", im["fuif_path"], "1"], stdout=subprocess.PIPE, text=True) for im in imgs]
            outs = [p.communicate()[0] for p in procs]
            inner = max(json.loads(o.strip().splitlines()[-1])["total_s"] for o in outs)
        else:
            from oracle import pyoracle as po
            res = [None] * n_img

            def work(i):
                t = time.perf_counter()
                o = po.OracleImage.decode(imgs[i]["fuif"])
                o.undo_transforms(0)
                res[i] = time.perf_counter() - t
            ths = [threading.Thread(target=work, args=(i,)) for i in range(n_img)]
            [t.start() for t in ths]
            [t.join() for t in ths]
            inner = max(res)
        return inner, time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    times = [one_step()[0] for _ in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    value = mpix / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "decode Mpixels/s (bit-exact)", "value": value, "unit": "Mpx/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16",
        "data": "synthetic", "config": {"workload": f"{args.workload}: {wl.WORKLOADS[args.workload][7]}", "images": n_img},
        "cpu_baseline": {"value": value, "unit": "Mpx/s", "cores": min(n_img, os.cpu_count() or 1), "kind": "reference" if use_ref else "port",
                         "host_cores": os.cpu_count(),
                         "sample": f"{n_img} full decode(s) of the workload image per step, one single-threaded process per image, all started together"},
        "e2e": {"value": value, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from fuif_b200 import api, shard
    import bench_workloads as wl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fuif_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- workload: one image per GPU (weak scaling, no data-path collective: files are independent units)
    spec = wl.WORKLOADS[args.workload]
    if args.strong:                 # a fixed batch split over the ranks (BASELINE config 4: 64 images over 8 GPUs)
        units = shard.split_batch(spec[5], rank, world)
        scaling = "strong"
    elif args.distinct_images:      # weak scaling, a different image on every rank
        units = shard.units_for_rank(spec[5], rank)
        scaling = "weak"
    else:                           # weak scaling, every rank decodes its own copy of the same image(s): the per-GPU work is identical
        units = shard.replica_units(spec[5])
        scaling = "weak"
    n_per_gpu = len(units)
    if world > 1:           # rank 0 encodes / indexes whatever is missing for EVERY rank first (one set of host processes, one GPU decode); the others then find the cache
        all_units = sorted(set(u for r in range(world) for u in (shard.split_batch(spec[5], r, world) if args.strong else
                                                                  shard.units_for_rank(spec[5], r) if args.distinct_images else shard.replica_units(spec[5]))))
        if rank == 0:
            wl.prepare_images(args.workload, all_units, want_index=not args.no_index)
        dist.barrier()
    imgs = wl.prepare_images(args.workload, units, want_index=not args.no_index)
    w, h, c, maxval = spec[0], spec[1], spec[2], spec[3]
    mpix_rank = n_per_gpu * w * h / 1e6
    bps = 2 if maxval > 255 else 1

    # a real (non-default) stream: the library enqueues on it and the CUDA events below are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = api.Context(local_rank, stream.cuda_stream)

    # compressed files resident in HBM (value path) and in pinned host memory (e2e path)
    dev_bytes = [torch.frombuffer(bytearray(im["fuif"]), dtype=torch.uint8).cuda() for im in imgs]
    pin_bytes = [torch.frombuffer(bytearray(im["fuif"]), dtype=torch.uint8).pin_memory() for im in imgs]
    pin_out = [torch.empty((h, w, c), dtype=(torch.int16 if bps == 2 else torch.uint8)).pin_memory() for _ in imgs]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def step_value(ev=None, indexed=True):
        """decode + undo_transforms, inputs and outputs in HBM. ev: optional (start, mid, end) CUDA events."""
        use_idx = indexed and not args.no_index
        if ev:
            ev[0].record(stream)
        if len(imgs) == 1:
            out = [api.fuif_decode((dev_bytes[0].data_ptr(), dev_bytes[0].numel()), ctx=ctx, group_index=imgs[0]["index"] if use_idx else None)]
        else:       # a batch: every (image, group) is a stream of ONE launch
            out = api.fuif_decode_batch([(db.data_ptr(), db.numel()) for db in dev_bytes], ctx=ctx,
                                        group_indexes=[im["index"] for im in imgs] if use_idx else None)
        if ev:
            ev[1].record(stream)
        for o in out:
            o.undo_transforms(0)
        if ev:
            ev[2].record(stream)
        return out

    def step_e2e(indexed=True):
        use_idx = indexed and not args.no_index
        if len(imgs) == 1:
            arr = np.frombuffer(memoryview(pin_bytes[0].numpy()), dtype=np.uint8)
            api.decode_to_pixels(arr, ctx=ctx, group_index=imgs[0]["index"] if use_idx else None, out=pin_out[0].numpy())
            return
        res_ = api.fuif_decode_batch([np.frombuffer(memoryview(pb.numpy()), dtype=np.uint8) for pb in pin_bytes], ctx=ctx,
                                     group_indexes=[im["index"] for im in imgs] if use_idx else None)
        for o in res_:
            o.undo_transforms(0)
        for o, po_ in zip(res_, pin_out):
            ctx.check(ctx.lib.fb_image_download_interleaved(o._handle, c, bps, po_.numpy().ctypes.data), "fb_image_download_interleaved")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- bit-exactness of what is being timed (rank 0, against the input image: the workload is lossless for cfg1/2/4)
    res = step_value()
    torch.cuda.synchronize()
    lossless = args.workload in ("cfg1", "cfg2", "cfg4", "mid")
    spec_seed0 = spec[4] + units[0]
    exact = None
    if lossless:
        from fuif_b200.synth import synth_image
        exact = bool(np.array_equal(res[0].pixels(), synth_image(w, h, c, maxval, spec_seed0)))
        if not exact:
            raise SystemExit("decoded pixels differ from the input image: refusing to report a number")
    del res

    for _ in range(max(0, args.warmup - 1)):
        r = step_value(); del r
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, CUDA events on the launching stream, L2 flushed between steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    launches0 = ctx.launches
    ctx.enable_kernel_timing(True)      # one CUDA event after every launch of the library (on its stream)
    ctx.timing_report()
    kernel_us = {}
    launch_seq = []         # per step: [(name, us, bytes)] in launch order
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        r = step_value(evs[k])
        del r
        rep = ctx.timing_report()                           # synchronises the stream
        launch_seq.append(rep)
        for name, us, nbytes in rep:
            kernel_us.setdefault(name, []).append((us, nbytes))
    ctx.enable_kernel_timing(False)
    torch.cuda.synchronize()
    launches = ctx.launches - launches0
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop()
    dec_ms = [e[0].elapsed_time(e[1]) for e in evs]
    chain_ms = [e[1].elapsed_time(e[2]) for e in evs]
    step_ms = [e[0].elapsed_time(e[2]) for e in evs]
    total_ms, total_units = shard.reduce_step_time(sum(step_ms), len(units), device="cuda")
    ms_per_step = total_ms / args.steps
    value = total_units * (w * h / 1e6) / (ms_per_step / 1e3)
    # every rank's own stage times (the max over ranks above hides which rank / stage is the slow one)
    mine = torch.tensor([sum(dec_ms) / len(dec_ms), sum(chain_ms) / len(chain_ms), sum(step_ms) / len(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    per_rank = [{"rank": i, "entropy_ms": float(t[0]), "transform_chain_ms": float(t[1]), "step_ms": float(t[2])} for i, t in enumerate(allr)]

    # ---- e2e: host buffers through the C-ABI convenience call
    for _ in range(min(args.warmup, 1)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_max, _ = shard.reduce_step_time(e2e_s * 1e3, len(units), device="cuda")
    e2e_value = total_units * (w * h / 1e6) / (e2e_max / 1e3 / args.steps)
    if lossless:
        from fuif_b200.synth import synth_image
        got = pin_out[0].numpy()
        if bps == 2:
            got = got.view(">u2")
        assert np.array_equal(got.astype(np.int32), synth_image(w, h, c, maxval, spec_seed0)), "e2e pixels differ"

    # ---- the same e2e call with the HOST-THREADS entropy backend (FB_OPT_ENTROPY_BACKEND = FB_ENTROPY_HOST): the serial
    # MANIAC coder of each channel group runs on a CPU thread, the planes go to HBM through pinned staging and the transform chain
    # runs on the GPU as before.  A second number next to `e2e`, not a replacement: `value` / `e2e` stay the all-GPU path.
    alt_entropy = {}
    if args.host_entropy_steps > 0:
        nthreads = args.host_threads or max(1, 4 * (os.cpu_count() or 1) // world)      # the library's default, shared out over the ranks
        for backend in ["host"]:
            ctx.set_entropy_backend(backend, nthreads)
            step_e2e()          # warm-up: the pinned staging is allocated once per context
            launches0 = ctx.launches
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.host_entropy_steps):
                step_e2e()
            torch.cuda.synchronize()
            he_ms, _ = shard.reduce_step_time((time.perf_counter() - t0) * 1e3, len(units), device="cuda")
            if lossless:
                for k_img, po_ in enumerate(pin_out):
                    got = po_.numpy()
                    if bps == 2:
                        got = got.view(">u2")
                    assert np.array_equal(got.astype(np.int32), synth_image(w, h, c, maxval, spec[4] + units[k_img])), f"{backend}-entropy e2e pixels differ (image {k_img})"
            alt_entropy[backend] = {
                "value": total_units * (w * h / 1e6) / (he_ms / 1e3 / args.host_entropy_steps), "unit": "Mpx/s",
                "ms_per_step": he_ms / args.host_entropy_steps, "steps": args.host_entropy_steps, "threads_per_gpu": ctx.host_threads_used,
                "host_cores": os.cpu_count(), "gpu_launches_per_step": (ctx.launches - launches0) // args.host_entropy_steps,
                "what": "same call as e2e with the entropy stage on host threads (one channel group per thread), transform chain on the GPU"}
            if not args.no_index and args.no_index_steps > 0:
                # ... and without the sidecar: what the reference arm gets too (one stream per file; a large plane may add a look-ahead
                # helper thread)
                barrier()
                t0 = time.perf_counter()
                step_e2e(indexed=False)
                torch.cuda.synchronize()
                hn_ms, _ = shard.reduce_step_time((time.perf_counter() - t0) * 1e3, len(units), device="cuda")
                if lossless:
                    got = pin_out[0].numpy()
                    if bps == 2:
                        got = got.view(">u2")
                    assert np.array_equal(got.astype(np.int32), synth_image(w, h, c, maxval, spec_seed0)), "un-indexed host-entropy e2e pixels differ"
                alt_entropy[backend]["no_index"] = {"value": total_units * (w * h / 1e6) / (hn_ms / 1e3), "unit": "Mpx/s", "ms_per_step": hn_ms, "steps": 1,
                                                    "worker_threads_per_gpu": ctx.host_threads_used}
        ctx.set_entropy_backend("gpu")
    host_entropy = alt_entropy.get("host")

    # ---- the same workload WITHOUT the group-offset sidecar (one stream per file, as the bare format dictates): a bounded sample
    no_index = None
    if not args.no_index and args.no_index_steps > 0:
        ev_n = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.no_index_steps)]
        barrier()
        for k in range(args.no_index_steps):
            flush.zero_()
            r = step_value(ev_n[k], indexed=False)
            del r
        torch.cuda.synchronize()
        ni_ms, _ = shard.reduce_step_time(sum(e[0].elapsed_time(e[2]) for e in ev_n), len(units), device="cuda")
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.no_index_steps):
            step_e2e(indexed=False)
        torch.cuda.synchronize()
        ni_e2e, _ = shard.reduce_step_time((time.perf_counter() - t0) * 1e3, len(units), device="cuda")
        if lossless:
            got = pin_out[0].numpy()
            if bps == 2:
                got = got.view(">u2")
            assert np.array_equal(got.astype(np.int32), synth_image(w, h, c, maxval, spec_seed0)), "un-indexed e2e pixels differ"
        no_index = {"value": total_units * (w * h / 1e6) / (ni_ms / 1e3 / args.no_index_steps),
                    "e2e": total_units * (w * h / 1e6) / (ni_e2e / 1e3 / args.no_index_steps), "unit": "Mpx/s", "steps": args.no_index_steps,
                    "ms_per_step": ni_ms / args.no_index_steps,
                    "what": "same workload, no sidecar: the channel groups of a file decode back to back on one SM (group offsets are not in the bitstream)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the transform chain, from the events inside the timed steps
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak = 6650.0; peak_src = "fallback 6.65 TB/s (B200_PROFILING.md)"
    alg_bytes = 4.0 * w * h * c * n_per_gpu
    chain_mean_ms = sum(chain_ms) / len(chain_ms)
    # The chain is ~0.4 ms of launches issued right after the host thread slept through a multi-second entropy stage: one late
    # launch (the host thread rescheduled, a cold core) puts idle time between two event marks that is not kernel time.  The
    # roofline therefore uses the MEDIAN over the timed steps, per launch and for the chain; the means are reported next to it.
    chain_med_ms = float(np.median(chain_ms))
    chain_gbs = alg_bytes / (chain_med_ms / 1e3) / 1e9
    per_kernel = {name: {"launches_per_step": len(v) / args.steps, "mean_us": sum(u for u, _ in v) / len(v),
                         "us_per_step": sum(u for u, _ in v) / args.steps, "algorithmic_bytes": v[0][1]} for name, v in kernel_us.items()}
    # dominant kernel of the chain = the single launch of the inverse transform chain with the largest mean duration (position by
    # position in the launch sequence of a step; the MANIAC launch is not on a bandwidth roofline)
    KERNEL_DOC = {
        "k_pk_hsq(ycocg)": "final horizontal unsqueeze step of Co and Cg fused with inverse YCoCg + clamp, packed int16x2, TMA-fed tiles (reads the "
                           "Co/Cg averages and residuals + Y, writes R, G, B): its algorithmic bytes equal the image's 4*W*H*C",
        "k_pk_hsq": "one horizontal unsqueeze step (packed int16x2, TMA-fed tiles), all planes of the step in one launch",
        "k_pk_vsq": "one vertical unsqueeze step (packed int16x2, 16-byte coalesced accesses), all planes of the step in one launch",
        "k_inv_hsq_direct(ycocg)": "final horizontal unsqueeze step of Co and Cg fused with inverse YCoCg + clamp (32-bit kernel)",
        "k_inv_hsq_direct": "one horizontal unsqueeze step, all planes of the step in one launch (32-bit kernel)",
        "k_inv_vsq_direct": "one vertical unsqueeze step, all planes of the step in one launch (32-bit kernel)",
        "k_fq_tiles(last)": "final unsqueeze steps of every plane + inverse YCoCg + clamp, one fused tile kernel",
        "k_idct_ycbcr": "dequantise + 8x8 inverse DCT + inverse YCbCr + clamp of all three components in one launch",
    }
    nl = min(len(r) for r in launch_seq) if launch_seq else 0
    same_seq = nl > 0 and all(len(r) == nl and [x[0] for x in r] == [x[0] for x in launch_seq[0]] for r in launch_seq)
    launches_ranked = []
    if same_seq:
        for pos in range(nl):
            name = launch_seq[0][pos][0]
            us = float(np.median([r[pos][1] for r in launch_seq]))
            launches_ranked.append({"pos": pos, "name": name, "us": us, "us_mean": sum(r[pos][1] for r in launch_seq) / len(launch_seq),
                                    "algorithmic_bytes": launch_seq[0][pos][2]})
    chain_launches = [x for x in launches_ranked if "maniac" not in x["name"]]
    cand = [x for x in chain_launches if x["algorithmic_bytes"] > 0]
    dom = max(cand, key=lambda x: x["us"]) if cand else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and dom:
        traffic = json.load(open(tpath)).get(args.workload, {}).get(dom["name"])
    if dom:
        achieved = dom["algorithmic_bytes"] / (dom["us"] * 1e-6) / 1e9
        roofline = {"bound": "hbm", "kernel": f"{dom['name']}: {KERNEL_DOC.get(dom['name'], '')}",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "algorithmic_bytes": dom["algorithmic_bytes"], "us": dom["us"], "us_mean": dom["us_mean"],
                    "stat": "median over the timed steps of the per-launch CUDA-event time (mean next to it)",
                    "share_of_chain": dom["us"] / (chain_med_ms * 1e3), "peak_source": peak_src,
                    "chain": {"what": "whole Image::undo_transforms (all launches), 4*W*H*C algorithmic bytes", "ms": chain_med_ms, "ms_mean": chain_mean_ms,
                              "achieved": chain_gbs, "frac": chain_gbs / peak, "launches": len(chain_launches)},
                    "chain_launches": [{"name": x["name"], "us": round(x["us"], 2), "MB": round(x["algorithmic_bytes"] / 1e6, 2)} for x in chain_launches]}
    else:       # no accounted launch: the whole chain
        roofline = {"bound": "hbm", "kernel": "inverse transform chain (Image::undo_transforms, all launches)",
                    "achieved": chain_gbs, "peak": peak, "unit": "GB/s", "frac": chain_gbs / peak, "traffic": None,
                    "algorithmic_bytes": alg_bytes, "ms": chain_med_ms, "ms_mean": chain_mean_ms, "peak_source": peak_src}

    # ---- reference CPU decoder on this box's host cores (1 core: it is single-threaded), one bounded sample
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        if wl.have_ref_driver():
            import tempfile
            from fuif_b200.synth import read_pnm
            with tempfile.TemporaryDirectory() as td:
                pnm = os.path.join(td, "ref.pnm")
                tt = wl.reference_decode_seconds(imgs[0]["fuif_path"], pnm)
                ref_pix, _ = read_pnm(pnm)
            got = pin_out[0].numpy()
            if bps == 2:
                got = got.view(">u2")
            exact_vs_ref = bool(np.array_equal(got.astype(np.int32), ref_pix))      # the reference is the checker here
            if not exact_vs_ref:
                raise SystemExit("GPU pixels differ from the reference decoder's pixels: refusing to report a number")
            cpu = {"value": w * h / 1e6 / tt["total_s"], "unit": "Mpx/s", "cores": 1, "kind": "reference", "gpu_pixels_equal_reference": exact_vs_ref,
                   "sample": "one full decode (fuif_decode_file + undo_transforms) of the first workload image by oracle/_ref/ref_driver",
                   "entropy_s": tt["entropy_s"], "chain_s": tt["chain_s"]}
        else:
            from oracle import pyoracle as po
            t0 = time.perf_counter()
            o = po.OracleImage.decode(imgs[0]["fuif"]); t1 = time.perf_counter()
            o.undo_transforms(0); t2 = time.perf_counter()
            cpu = {"value": w * h / 1e6 / (t2 - t0), "unit": "Mpx/s", "cores": 1, "kind": "port",
                   "sample": "one full decode of the first workload image by oracle/libfuif_oracle.so", "entropy_s": t1 - t0, "chain_s": t2 - t1}

    line = {
        "metric": "decode Mpixels/s (bit-exact)", "value": value, "unit": "Mpx/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec[7]}", "images": int(total_units), "images_per_gpu": n_per_gpu,
                   "width": w, "height": h, "channels": c,
                   "group_index_sidecar": not args.no_index, "group_index_source": None if args.no_index else imgs[0].get("index_source"),
                   "l2": "256 MiB buffer written between timed steps",
                   "parallelism": (f"fixed batch of {spec[5]} images split over {world} GPU(s)" if scaling == "strong" else
                                   f"{n_per_gpu} image(s) per GPU x{world}" + ("" if args.distinct_images else " (replicas of the same image(s))")),
                   "bit_exact_checked": exact},
        "value_no_index": no_index["value"] if no_index else None, "e2e_no_index": no_index["e2e"] if no_index else None, "no_index": no_index,
        "e2e": {"value": e2e_value, "unit": "Mpx/s", "h2d_bytes_per_step": int(sum(len(im["fuif"]) for im in imgs)),
                "d2h_bytes_per_step": int(n_per_gpu * w * h * c * bps)},
        "e2e_host_entropy": host_entropy, "e2e_host_entropy_no_index": (host_entropy or {}).get("no_index", {}).get("value"),
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "per_rank": per_rank,
        "stages": {"entropy_ms": sum(dec_ms) / len(dec_ms), "transform_chain_ms": chain_mean_ms, "transform_chain_ms_median": chain_med_ms, "wall_s_timed_region": t_wall,
                   "transform_chain_mpx_s": mpix_rank / (chain_mean_ms / 1e3), "kernels": per_kernel,
                   "unsqueeze_repaired_segments": ctx.pk_repaired, "unsqueeze_range_flagged_segments": ctx.pk_range_flagged},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
