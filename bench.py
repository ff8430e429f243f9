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
    n_img = max(1, args.gpus) * wl.WORKLOADS[args.workload][5]
    imgs = wl.prepare_images(args.workload, range(n_img), want_index=False)
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
        "cpu_baseline": {"value": value, "unit": "Mpx/s", "cores": n_img, "kind": "reference" if use_ref else "port",
                         "sample": f"{n_img} full decode(s) of the workload image per step, one single-threaded process per image"},
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
    n_per_gpu = spec[5]
    units = shard.units_for_rank(n_per_gpu, rank)
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

    def step_value(ev=None):
        """decode + undo_transforms, inputs and outputs in HBM. ev: optional (start, mid, end) CUDA events."""
        if ev:
            ev[0].record(stream)
        if len(imgs) == 1:
            out = [api.fuif_decode((dev_bytes[0].data_ptr(), dev_bytes[0].numel()), ctx=ctx, group_index=imgs[0]["index"])]
        else:       # a batch: every (image, group) is a stream of ONE launch
            out = api.fuif_decode_batch([(db.data_ptr(), db.numel()) for db in dev_bytes], ctx=ctx,
                                        group_indexes=None if args.no_index else [im["index"] for im in imgs])
        if ev:
            ev[1].record(stream)
        for o in out:
            o.undo_transforms(0)
        if ev:
            ev[2].record(stream)
        return out

    def step_e2e():
        if len(imgs) == 1:
            arr = np.frombuffer(memoryview(pin_bytes[0].numpy()), dtype=np.uint8)
            api.decode_to_pixels(arr, ctx=ctx, group_index=imgs[0]["index"], out=pin_out[0].numpy())
            return
        res_ = api.fuif_decode_batch([np.frombuffer(memoryview(pb.numpy()), dtype=np.uint8) for pb in pin_bytes], ctx=ctx,
                                     group_indexes=None if args.no_index else [im["index"] for im in imgs])
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
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        r = step_value(evs[k])
        del r
        for name, us, nbytes in ctx.timing_report():       # synchronises the stream
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
    chain_gbs = alg_bytes / (chain_mean_ms / 1e3) / 1e9
    per_kernel = {name: {"launches_per_step": len(v) / args.steps, "mean_us": sum(u for u, _ in v) / len(v),
                         "us_per_step": sum(u for u, _ in v) / args.steps, "algorithmic_bytes": v[0][1]} for name, v in kernel_us.items()}
    # dominant kernel of the chain = the launch of the library with the largest time per step among those whose
    # algorithmic bytes the library accounts (the MANIAC launch is not on a bandwidth roofline)
    KERNEL_DOC = {
        "k_inv_hsq_direct(ycocg)": "final horizontal unsqueeze step of Co and Cg fused with inverse YCoCg + clamp (reads Co/Cg averages and "
                                   "residuals + Y, writes R, G, B): its algorithmic bytes equal the image's 4*W*H*C",
        "k_inv_hsq_direct": "one horizontal unsqueeze step, all planes of the step in one launch",
        "k_inv_vsq_direct": "one vertical unsqueeze step, all planes of the step in one launch",
        "k_fq_tiles(last)": "final unsqueeze steps of every plane + inverse YCoCg + clamp, one fused tile kernel",
    }
    cand = [(sum(u for u, _ in v) / args.steps, name) for name, v in kernel_us.items() if v[0][1] and v[0][1] > 0 and "maniac" not in name]
    dom_name = max(cand)[1] if cand else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and dom_name:
        traffic = json.load(open(tpath)).get(args.workload, {}).get(dom_name)
    dom = kernel_us.get(dom_name) if dom_name else None
    if dom and dom[0][1] > 0:
        # several launches of that name per step (one per squeeze step): the roofline is quoted for the largest one
        per_step = len(dom) // args.steps
        big = max(range(per_step), key=lambda i: dom[i][1]) if per_step > 1 else 0
        sel = [dom[k * per_step + big] for k in range(args.steps)] if per_step >= 1 else dom
        us = sum(u for u, _ in sel) / len(sel)
        kbytes = sel[0][1]
        achieved = kbytes / (us * 1e-6) / 1e9
        roofline = {"bound": "hbm", "kernel": f"{dom_name}: {KERNEL_DOC.get(dom_name, '')}",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "algorithmic_bytes": kbytes, "us": us, "share_of_chain": us / (chain_mean_ms * 1e3),
                    "peak_source": peak_src,
                    "chain": {"what": "whole Image::undo_transforms (all launches), 4*W*H*C algorithmic bytes", "ms": chain_mean_ms,
                              "achieved": chain_gbs, "frac": chain_gbs / peak}}
    else:       # chains without an accounted Squeeze launch (e.g. the DCT chain): the whole chain
        roofline = {"bound": "hbm", "kernel": "inverse transform chain (Image::undo_transforms, all launches)",
                    "achieved": chain_gbs, "peak": peak, "unit": "GB/s", "frac": chain_gbs / peak, "traffic": None,
                    "algorithmic_bytes": alg_bytes, "ms": chain_mean_ms, "peak_source": peak_src}

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
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec[7]}", "images_per_gpu": n_per_gpu, "width": w, "height": h, "channels": c,
                   "group_index_sidecar": not args.no_index, "l2": "256 MiB buffer written between timed steps", "parallelism": f"one image per GPU x{world}",
                   "bit_exact_checked": exact},
        "e2e": {"value": e2e_value, "unit": "Mpx/s", "h2d_bytes_per_step": int(sum(len(im["fuif"]) for im in imgs)),
                "d2h_bytes_per_step": int(n_per_gpu * w * h * c * bps)},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stages": {"entropy_ms": sum(dec_ms) / len(dec_ms), "transform_chain_ms": chain_mean_ms, "wall_s_timed_region": t_wall,
                   "transform_chain_mpx_s": mpix_rank / (chain_mean_ms / 1e3), "kernels": per_kernel,
                   "unsqueeze_repaired_tiles": ctx.repaired_tiles, "unsqueeze_serial_fallbacks": ctx.fallbacks},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
