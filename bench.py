#!/usr/bin/env python
"""bench.py -- Ascent `volume` plot hot path on B200: Mrays/s, frames/s, composite ms/frame.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c3]

One "step" = one frame of the hot path.
  N == 1 : BASELINE config 2 -- braid uniform 512^3 f32, one domain, 1920x1080, default camera,
           samples = 100, path A of vtkh::VolumeRenderer (clear canvas, trace [K1-K7], Image::Init
           quantise, ImageToCanvas), field resident in HBM.
  N  > 1 : BASELINE config 3 -- braid 1023^3 split into 8 blocks of 512^3 spread over N ranks
           (8/N blocks per rank), 3840x2160, sort-last render + composite to rank 0 (strong
           scaling; N == 8 is path A/direct-send, N < 8 path B/partials like the reference).
`value` = primary rays of the final image (W*H) x frames / time, device-timed (CUDA events on the
launching stream, max over ranks), inputs resident.  `e2e` = same frame through the host-buffer
C ABI a vtk-h caller uses: field uploaded from pinned host memory, canvas read back, every step.
`--impl reference` times the CPU oracle (the restatement of the reference's OpenMP path, see
oracle/) on the same workload with every host thread.
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

RAMP_TF = {"name": "cool to warm", "control_points": [
    {"type": "alpha", "position": 0., "alpha": 0.}, {"type": "alpha", "position": 1., "alpha": 1.}]}
SAMPLES = 100  # the reference's default (VolumeRenderer.cpp:406); --samples overrides (P1 = 887)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- workloads
def workload_c2():
    return dict(name="c2: braid uniform 512^3 f32, 1 domain, 1920x1080, default camera, samples=100",
                n_block=512, per_axis=1, W=1920, H=1080, key="c2")


def workload_c3():
    return dict(name="c3: braid 1023^3 as 8 blocks of 512^3 f32, 3840x2160, default camera, "
                     "samples=100, sort-last composite to rank 0",
                n_block=512, per_axis=2, W=3840, H=2160, key="c3")


def workload_c4():
    return dict(name="c4: braid rectilinear 768^3 f32 (axes warped, power 1.5), 64-view cinema orbit "
                     "(phi=8, theta=8), 1024x1024, samples=100; one step = all 64 views",
                n_block=768, per_axis=1, W=1024, H=1024, key="c4", rectilinear=True, views=64)


def workload_c5():
    return dict(name="c5: 512 uniform domains of 128^3 f32 (global 1017^3), 4096x4096, default camera, "
                     "samples=100, partial compositing (ray layers)",
                n_block=128, per_axis=8, W=4096, H=4096, key="c5")


def block_layout(wl):
    """origin/spacing/start index of every block of the workload (same as
    ascent_b200.datasets.braid_uniform_blocks)."""
    nb, b = wl["n_block"], wl["per_axis"]
    g = b * (nb - 1) + 1
    sp = 20.0 / (g - 1)
    out = []
    for kz in range(b):
        for jy in range(b):
            for ix in range(b):
                st = (ix * (nb - 1), jy * (nb - 1), kz * (nb - 1))
                out.append(dict(dims=(nb,) * 3, start=st, glob=(g,) * 3, spacing=[sp] * 3,
                                origin=[-10.0 + st[0] * sp, -10.0 + st[1] * sp, -10.0 + st[2] * sp]))
    return out


def scene_params(wl, blocks, rng_minmax):
    """Host-side driver values of VolumeRenderer::PreExecute for the workload."""
    from ascent_b200 import _lib, camera, color_table, datasets
    bl = [datasets.domain_bounds(dict(kind="uniform", dims=b["dims"], origin=b["origin"],
                                      spacing=b["spacing"])) for b in blocks]
    gb = datasets.union_bounds(bl)
    cam = camera.Camera()
    cam.reset_to_bounds(gb)
    cams = None
    if wl.get("views"):
        cams = [c.to_struct() for c in camera.cinema_cameras(gb, *camera.cinema_angles(8, 8))]
    lut = color_table.parse_color_table(RAMP_TF).corrected_opacity(SAMPLES).lut()
    return dict(bounds=bl, gb=gb, cam=cam.to_struct(), cams=cams, lut=lut,
                sample_dist=_lib.sample_distance(gb, SAMPLES), rmin=rng_minmax[0], rmax=rng_minmax[1])


# ----------------------------------------------------------------------------- CPU arm
def run_reference(args, wl):
    """The reference's CPU path (oracle restatement, all host threads) on the same workload."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ascent_b200 import datasets
    from oracle import oracle as O
    blocks = block_layout(wl)
    W, H = wl["W"], wl["H"]
    t0 = time.time()
    doms = []
    for b in blocks:
        f = datasets.braid_values(*b["dims"], *b["start"], *b["glob"], dtype=np.float32)
        doms.append(dict(kind="uniform", dims=b["dims"], origin=b["origin"], spacing=b["spacing"], field=f))
    rmin = min(float(d["field"].min()) for d in doms)
    rmax = max(float(d["field"].max()) for d in doms)
    sp = scene_params(wl, blocks, (rmin, rmax))
    cam = O.Camera.from_buffer_copy(bytes(sp["cam"]))
    oblocks = [O.OracleBlock(d["dims"], d["field"], origin=d["origin"], spacing=d["spacing"]) for d in doms]
    gen_s = time.time() - t0

    def frame():
        rgba, depth = O.new_canvas(W, H)
        if len(oblocks) == 1:
            O.render_to_canvas(oblocks[0], cam, W, H, sp["lut"], sp["sample_dist"], rmin, rmax, rgba, depth)
            u8, d = O.image_init(rgba, depth, 0)
            O.image_to_canvas(u8, d)
        else:
            pl = [O.render_partials(ob, cam, W, H, sp["lut"], sp["sample_dist"], rmin, rmax, depth)
                  for ob in oblocks]
            O.partials_to_canvas(O.composite_partials(pl), cam, W, H, rgba, depth)

    for _ in range(args.warmup):
        frame()
    t0 = time.time()
    for _ in range(args.steps):
        frame()
    dt = (time.time() - t0) / args.steps
    val = W * H / dt / 1e6
    line = {"impl": "reference", "metric": "volume_render_mrays_per_s", "value": val, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "image": [W, H], "samples": SAMPLES},
            "frames_per_s": 1.0 / dt,
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": O.num_threads(), "kind": "port",
                             "sample": "whole frame(s) of the workload, oracle/raycast_oracle.c + "
                                       "composite_oracle.c with OpenMP (VTK-m/OpenMP reference is not "
                                       "buildable here); field generation %.1fs excluded" % gen_s},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm, N == 1
def run_single(args, wl):
    import torch
    from ascent_b200 import _lib
    torch.cuda.set_device(0)
    ctx = _lib.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    blocks = block_layout(wl)
    W, H = wl["W"], wl["H"]
    nvox = int(np.prod(blocks[0]["dims"]))
    fields = []
    with torch.cuda.stream(stream):
        for i, b in enumerate(blocks):
            t = torch.empty(nvox, dtype=torch.float32, device="cuda")
            ctx.synth_braid_dev(t.data_ptr(), _lib.VR_F32, b["dims"], b["start"], b["glob"])
            fields.append(t)
        stream.synchronize()
        rmin = min(float(t.min()) for t in fields)
        rmax = max(float(t.max()) for t in fields)
    sp = scene_params(wl, blocks, (rmin, rmax))
    ctx.set_tf(sp["lut"])
    for i, b in enumerate(blocks):
        if wl.get("rectilinear"):
            from ascent_b200 import datasets
            axes = [datasets.warp_axis(n, 1.5) for n in b["dims"]]
            ctx.block_rectilinear(i, b["dims"], axes, None, device_ptr=fields[i].data_ptr(), dtype=_lib.VR_F32)
        else:
            ctx.block_uniform(i, b["dims"], b["origin"], b["spacing"], None, device_ptr=fields[i].data_ptr(),
                              dtype=_lib.VR_F32)
    cam = sp["cam"]
    multi = len(blocks) > 1
    views = sp["cams"] or [cam]

    def frame(ev=None):
        if not multi:
            # RenderOneDomainPerRank on a cleared canvas: Canvas::Clear + RenderCells + Image::Init +
            # ImageToCanvas in one launch (vr_trace_to_image); the cinema workload renders all its
            # views back to back (the reference batches <= 10 renders per Update, Scene.cpp:133-149)
            if ev:
                ev[0].record(stream)
            for v in views:
                ctx.trace_to_image(0, v, W, H, sp["sample_dist"], rmin, rmax, write_canvas=True)
            if ev:
                ev[1].record(stream)
        else:
            # RenderMultipleDomainsPerRank on dense ray layers: one sampler launch per block (one ABI
            # call for the whole loop, consecutive blocks overlapped on side streams), then
            # PartialCompositor::composite + partials_to_canvas over a cleared canvas in ONE kernel
            ctx.layers_begin(W, H)
            if ev:
                ev[0].record(stream)
            ctx.trace_blocks_to_layers(list(range(len(blocks))), cam, sp["sample_dist"], rmin, rmax, False)
            if ev:
                ev[1].record(stream)
            ctx.layers_composite_to_canvas(cam, canvas_is_clear=True)

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            frame()
        stream.synchronize()
        l0 = ctx.kernel_launches()
        clocks = ClockSampler(0)
        clocks.start()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t_begin.record(stream)
        for k in range(args.steps):
            frame(evs[k])
        t_end.record(stream)
        torch.cuda.synchronize()
        clk = clocks.stop()
        launches = ctx.kernel_launches() - l0
        total_ms = t_begin.elapsed_time(t_end)
        trace_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    ms = total_ms / args.steps
    value = W * H * len(views) / (ms * 1e-3) / 1e6

    # ---- e2e: host buffers through the ABI a vtk-h caller uses, H2D + D2H inside the timed region
    e2e = None
    if not multi:
        host_field = torch.empty(nvox, dtype=torch.float32, pin_memory=True)
        host_field.copy_(fields[0])
        rgba_h = torch.zeros(H * W * 4, dtype=torch.float32, pin_memory=True)
        depth_h = torch.full((H * W,), 1.001, dtype=torch.float32, pin_memory=True)
        hf = host_field.numpy()
        hr, hd = rgba_h.numpy().reshape(-1, 4), depth_h.numpy()
        b = blocks[0]

        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

        def e2e_frame(mode):
            # publish: the simulation's field (pinned host memory), every step.
            #  "copy"   : the whole block goes host -> device (cudaMemcpyAsync) before the trace;
            #  "mapped" : the block is registered in place (VR_HOST_MAPPED), the sampler reads it over PCIe;
            #  "staged" : VR_HOST_STAGED -- a pre-pass of the sampler flags the 128-byte lines the rays of
            #             this view touch, a gather kernel pulls exactly those over PCIe, then the trace.
            # For the in-place modes L2 is flushed first so that no line of the previous step's
            # (identical) field can be served from cache.
            if mode != "copy":
                with torch.cuda.stream(stream):
                    flush.zero_()
            kw = dict(host_mapped=(mode == "mapped"), staged=(mode == "staged"))
            if wl.get("rectilinear"):
                ctx.block_rectilinear(100, b["dims"], axes, hf, **kw)
            else:
                ctx.block_uniform(100, b["dims"], b["origin"], b["spacing"], hf, **kw)
            # the volume-only scene of the config: the frame starts from a cleared canvas, so the
            # whole per-rank body is one launch; the result comes back as the float canvas
            for v in views:
                ctx.trace_to_image(100, v, W, H, sp["sample_dist"], rmin, rmax, write_canvas=True)
                ctx.canvas_download(W, H, hr, hd)

        n_e2e = max(3, min(args.steps, 10))
        modes, moved = {}, {}
        for name in ("copy", "mapped", "staged"):
            if name == "mapped" and len(views) > 8:
                continue  # many views per publish: in-place sampling re-reads the block for every view
            for _ in range(2):
                e2e_frame(name)
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                e2e_frame(name)
            modes[name] = (time.perf_counter() - t0) / n_e2e
            if name == "staged":
                moved[name] = ctx.block_staged_bytes(100)
        best = min(modes, key=modes.get)
        dt = modes[best]
        # the same frame through the canvas-in/canvas-out form vtk-h's RenderCells seam needs when
        # opaque geometry is already on the canvas (upload + K2 depth clamp + blend over + download)
        ctx.block_uniform(100, b["dims"], b["origin"], b["spacing"], hf) if not wl.get("rectilinear") else \
            ctx.block_rectilinear(100, b["dims"], axes, hf)
        t0 = time.perf_counter()
        for _ in range(3):
            hr.fill(0.0)
            hd.fill(1.001)
            ctx.render_image(100, cam, W, H, sp["sample_dist"], rmin, rmax, hr, hd)
        dt_inout = (time.perf_counter() - t0) / 3
        touched = None
        tpj = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpj):
            touched = (json.load(open(tpj)).get(wl["key"]) or {}).get("dram_read_bytes")
        h2d = {"copy": nvox * 4, "mapped": touched or nvox * 4, "staged": moved.get("staged", nvox * 4)}
        e2e = {"value": W * H * len(views) / dt / 1e6, "unit": "Mrays/s", "ms_per_step": dt * 1e3,
               "mode": best,
               "h2d_bytes_per_step": h2d[best],
               "d2h_bytes_per_step": W * H * 20 * len(views),
               "modes_ms_per_step": {k: v * 1e3 for k, v in modes.items()},
               "modes_h2d_bytes": {k: h2d[k] for k in modes},
               "field_bytes": nvox * 4,
               "what": "per step: publish the pinned host field + vr_trace_to_image + vr_canvas_download(host "
                       "canvas) per view, all through the C ABI with host buffers.  copy: VR_HOST, cudaMemcpyAsync "
                       "of the whole block; staged: VR_HOST_STAGED, sampler pre-pass flags the 128-byte lines the "
                       "view's rays touch and a gather kernel pulls exactly those across PCIe (h2d bytes counted "
                       "by the library, L2 flushed every step); mapped: VR_HOST_MAPPED, sampled in place (h2d = "
                       "DRAM-read bytes of the ncu capture).  The fastest mode is the headline; images are "
                       "bit-identical in all three (tests/test_gpu_parity.py)",
               "render_image_canvas_inout_ms": dt_inout * 1e3}
        ctx.block_free(100)

    # ---- roofline of the dominant kernel (trace), algorithmic bytes per launch
    pk, pk_src = peaks()
    n_launch = len(blocks)
    # N_vox*4 + W*H*20 per frame (the fused kernel really writes W*H*28: RGBA8 + depth
    # image and the float canvas; the extra 8 B/pixel are not claimed)
    alg_bytes = nvox * 4 * len(blocks) + W * H * 20
    achieved = alg_bytes / (trace_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp)).get(wl["key"])
        if t:
            traffic, traffic_src = t["traffic_bytes"] * n_launch, t["source"]
    roof = {"bound": "hbm", "kernel": "trace_kernel (sampler.cu)", "achieved": achieved,
            "peak": pk["hbm_gbs"], "peak_source": pk_src, "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
            "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_frame": alg_bytes, "kernel_ms_per_frame": trace_ms,
            "launches_per_frame": n_launch}

    # ---- CPU baseline beside it: bounded sample of the same workload (rank 0, N == 1)
    cpu = cpu_baseline_sample(wl, blocks, fields, sp, rmin, rmax) if not args.no_cpu else None

    line = {"metric": "volume_render_mrays_per_s", "value": value, "unit": "Mrays/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "image": [W, H], "samples": SAMPLES,
                       "path": "B (ray layers)" if multi else "A (image)",
                       "l2": "inputs (%.0f MB field) larger than the 126 MB L2" % (nvox * 4 * len(blocks) / 1e6)},
            "frames_per_s": 1e3 / ms, "composite_ms_per_frame": ms - trace_ms,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line))
    ctx.close()


def cpu_baseline_sample(wl, blocks, fields, sp, rmin, rmax):
    """The oracle port of the reference's CPU path on the host cores, bounded to ~10 s: whole frames
    of the same workload (for the 64-view cinema workload: as many views as fit the budget)."""
    from ascent_b200 import datasets
    from oracle import oracle as O
    W, H = wl["W"], wl["H"]
    cams = [O.Camera.from_buffer_copy(bytes(c)) for c in (sp["cams"] or [sp["cam"]])]
    if wl.get("rectilinear"):
        obs = [O.OracleBlock(b["dims"], f.cpu().numpy(), axes=[datasets.warp_axis(n, 1.5) for n in b["dims"]])
               for b, f in zip(blocks, fields)]
    else:
        obs = [O.OracleBlock(b["dims"], f.cpu().numpy(), origin=b["origin"], spacing=b["spacing"])
               for b, f in zip(blocks, fields)]
    t0 = time.time()
    n = 0
    while True:
        cam = cams[n % len(cams)]
        rgba, depth = O.new_canvas(W, H)
        if len(obs) == 1:
            O.render_to_canvas(obs[0], cam, W, H, sp["lut"], sp["sample_dist"], rmin, rmax, rgba, depth)
            u8, d = O.image_init(rgba, depth, 0)
            O.image_to_canvas(u8, d)
        else:
            pl = [O.render_partials(ob, cam, W, H, sp["lut"], sp["sample_dist"], rmin, rmax, depth) for ob in obs]
            O.partials_to_canvas(O.composite_partials(pl), cam, W, H, rgba, depth)
        n += 1
        if time.time() - t0 > 10.0 or n >= 20:
            break
    dt = (time.time() - t0) / n
    return {"value": W * H / dt / 1e6, "unit": "Mrays/s", "cores": O.num_threads(), "kind": "port",
            "ms_per_frame": dt * 1e3,
            "sample": "%d whole frame(s)/view(s) of the same workload on the host cores (oracle port of the "
                      "reference's VTK-m/OpenMP algorithm; the reference itself is not buildable here)" % n}


def main():
    global SAMPLES
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--one-block-per-rank", action="store_true",
                    help="N > 1 diagnostics: rank r renders only block r of c3 (path A at any N); not a bench line")
    ap.add_argument("--nccl-baseline", action="store_true",
                    help="N > 1, image path: also time the NCCL all_to_all + fold + gather form of the exchange")
    ap.add_argument("--samples", type=int, default=None,
                    help="samples option of the volume plot (default 100; 887 = one sample per voxel on c2)")
    args = ap.parse_args()
    if args.samples:
        SAMPLES = args.samples
    wname = args.workload or ("c2" if args.gpus == 1 else "c3")
    wl = {"c2": workload_c2, "c3": workload_c3, "c4": workload_c4, "c5": workload_c5}[wname]()
    wl["name"] = wl["name"].replace("samples=100", "samples=%d" % SAMPLES)
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5
        args.warmup = min(args.warmup, 1)
        return run_reference(args, wl)
    if args.gpus == 1:
        return run_single(args, wl)
    from ascent_b200 import distributed
    return distributed.run_bench(args, wl, sys.modules[__name__])


if __name__ == "__main__":
    main()
