#!/usr/bin/env python
"""bench.py -- Ascent `volume` plot hot path on B200: Mrays/s, frames/s, composite ms/frame.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c3|c4|c5]

One "step" = one frame of the hot path.
  plain `--gpus 1`   : BASELINE config 2 -- braid uniform 512^3 f32, one domain, 1920x1080, default
           camera, samples = 100, path A of vtkh::VolumeRenderer (clear canvas, trace [K1-K7],
           Image::Init quantise, ImageToCanvas), field resident in HBM.  The line also carries
           `c3_n1`: config 3 on this one GPU (all 8 blocks, path B), the T(1) of the scaling curve.
  under torchrun     : BASELINE config 3 at every N, 1 included -- braid 1023^3 split into 8 blocks
           of 512^3 spread over N ranks (8/N blocks per rank), 3840x2160, sort-last render +
           composite to rank 0 (strong scaling; N == 8 is path A/direct-send, N < 8 path B/partials
           like the reference), so that a 1/2/4/8 sweep is one workload.
`value` = primary rays of the final image (W*H) x frames / time, device-timed (CUDA events on the
launching stream, max over ranks), inputs resident.  `e2e` = same frame through the host-buffer
C ABI a vtk-h caller uses: field uploaded from pinned host memory, canvas read back, every step.
`parity` = the timed configuration's final canvas against the CPU oracle on the same inputs,
computed outside the timed region.  `cpu_baseline` = the oracle port timed on the host cores.
`--impl reference` times the CPU oracle (the restatement of the reference's OpenMP path, see
oracle/) on the same workload with every host thread; it imports nothing of the product.
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
WARP_POWER = 1.5  # c4: axis x_i = -10 + 20 (i/(n-1))^1.5


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def under_torchrun():
    return "TORCHELASTIC_RUN_ID" in os.environ or all(k in os.environ for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))


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


def nvlink_counters(index):
    """(rx_bytes, tx_bytes) of GPU `index` since boot, summed over its NVLinks, from the NVML hardware counters
    (NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX/TX, payload bytes, reported in KiB) -- or None where NVML does not
    expose them.  Read before and after a run of frames, the difference is what the exchange really moved."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(index))
        ids = (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX)

        def read(scope):
            v = pynvml.nvmlDeviceGetFieldValues(h, [(ids[0], scope), (ids[1], scope)])
            if any(x.nvmlReturn != 0 for x in v):
                return None
            return [int(x.value.ullVal) * 1024 for x in v]

        tot = read(0xFFFFFFFF)  # all links at once
        if tot is None:        # ... or link by link
            tot = [0, 0]
            n_ok = 0
            for link in range(18):
                r = read(link)
                if r is not None:
                    tot[0] += r[0]
                    tot[1] += r[1]
                    n_ok += 1
            if n_ok == 0:
                return None
        return tuple(tot)
    except Exception:
        return None


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


WORKLOADS = {"c2": workload_c2, "c3": workload_c3, "c4": workload_c4, "c5": workload_c5}


def block_layout(wl):
    """origin/spacing/start index of every block of the workload (Conduit's braid on [-10,10]^3,
    blocks sharing one point layer)."""
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


def warp_axis(n, power=WARP_POWER):
    t = np.arange(n, dtype=np.float64) / float(n - 1)
    return -10.0 + 20.0 * t ** power


def block_bounds(wl, b):
    """coords.GetBounds() of a block: f64 arithmetic on the f32-valued origin/spacing (uniform) or the
    f64 axis end points (rectilinear)."""
    if wl.get("rectilinear"):
        out = []
        for n in b["dims"]:
            a = warp_axis(n)
            out += [float(a[0]), float(a[-1])]
        return np.array(out)
    out = []
    for a in range(3):
        o, s = float(np.float32(b["origin"][a])), float(np.float32(b["spacing"][a]))
        out += [o, o + s * float(b["dims"][a] - 1)]
    return np.array(out)


def union_bounds(bl):
    bl = np.asarray(bl, np.float64).reshape(-1, 6)
    out = np.zeros(6)
    out[0::2] = bl[:, 0::2].min(axis=0)
    out[1::2] = bl[:, 1::2].max(axis=0)
    return out


def image_path(n_blocks, n_gpus):
    """the reference's switch (VolumeRenderer.cpp:468-480): path A only when every rank holds one domain"""
    return n_blocks == n_gpus


def config_of(wl, n_gpus):
    """The `config` object of the JSON line -- the same dict from the B200 arm and the reference arm."""
    blocks = wl["per_axis"] ** 3
    per_gpu_mb = blocks * wl["n_block"] ** 3 * 4 / max(n_gpus, 1) / 1e6
    return {"workload": wl["name"], "image": [wl["W"], wl["H"]], "samples": SAMPLES, "blocks": blocks,
            "gpus": n_gpus,
            "path": ("A: one domain per rank -> uint8 images, visibility-ordered fold" if image_path(blocks, n_gpus)
                     else "B: several domains per rank -> float partials, depth-ordered fold"),
            "l2": "no flush: every frame reads %.0f MB of field per GPU, larger than the 126 MB L2" % per_gpu_mb}


# ----------------------------------------------------------------------------- oracle side (CPU)
class OracleScene:
    """The reference's CPU path for a workload, through oracle/ only (K0 camera, K8 table, V4 sample
    distance, K1-K7 trace, C1-C2 / P1-P4 compositing, V9/V10 canvas).  Used as the checker (`parity`),
    as the timed CPU baseline, and as the whole of `--impl reference`."""

    def __init__(self, wl, fields, n_gpus, rng=None):
        from oracle import oracle as O
        self.O = O
        self.wl = wl
        self.W, self.H = wl["W"], wl["H"]
        self.blocks = block_layout(wl)
        self.bounds = [block_bounds(wl, b) for b in self.blocks]
        self.gb = union_bounds(self.bounds)
        self.cam = O.camera_reset_to_bounds(self.gb)
        self.cams = None
        if wl.get("views"):
            ph, th = O.cinema_angles(8, 8)
            self.cams = [O.camera_cinema(self.gb, p, t) for p in ph for t in th]
        self.lut = O.parse_color_table(RAMP_TF).correct_opacity(SAMPLES).lut()
        self.sample_dist = O.sample_distance(self.gb, SAMPLES)
        if wl.get("rectilinear"):
            self.obs = [O.OracleBlock(b["dims"], f, axes=[warp_axis(n) for n in b["dims"]])
                        for b, f in zip(self.blocks, fields)]
        else:
            self.obs = [O.OracleBlock(b["dims"], f, origin=b["origin"], spacing=b["spacing"])
                        for b, f in zip(self.blocks, fields)]
        if rng is None:
            rng = (min(float(f.min()) for f in fields), max(float(f.max()) for f in fields))
        self.rmin, self.rmax = rng
        self.path_a = image_path(len(self.blocks), n_gpus)
        self.n_partials = 0

    def frame(self, view=0):
        """one frame: the final float canvas (rgba, depth) as rank 0 holds it"""
        O, W, H = self.O, self.W, self.H
        cam = self.cams[view % len(self.cams)] if self.cams else self.cam
        a = (cam, W, H, self.lut, self.sample_dist, self.rmin, self.rmax)
        if self.path_a:
            # RenderOneDomainPerRank + Composite: per-rank canvas -> Image::Init -> ordered uint8 fold
            # -> ImageToCanvas on rank 0
            layers, depths = [], []
            for ob in self.obs:
                rgba, depth = O.new_canvas(W, H)
                O.render_to_canvas(ob, *a, rgba, depth)
                u8, d = O.image_init(rgba, depth, 0)
                layers.append(u8)
                depths.append(d)
            if len(layers) == 1:
                out, od = layers[0], depths[0]
            else:
                order, _ = O.visibility_order(np.array(self.bounds), cam)
                out, od = O.ordered_composite(np.stack(layers), np.stack(depths), order)
            return O.image_to_canvas(out, od)
        # RenderMultipleDomainsPerRank: partials of every domain -> sort + fold -> partials_to_canvas
        rgba, depth = O.new_canvas(W, H)
        pl = [O.render_partials(ob, *a, depth) for ob in self.obs]
        self.n_partials = int(sum(p.size for p in pl))
        O.partials_to_canvas(O.composite_partials(pl), cam, W, H, rgba, depth)
        return rgba, depth

    def time_frames(self, n_max, budget_s):
        """whole frames (views) of the workload until the budget is spent"""
        t0 = time.time()
        n = 0
        while True:
            self.frame(n)
            n += 1
            if time.time() - t0 > budget_s or n >= n_max:
                break
        return (time.time() - t0) / n, n


def compare_canvas(mine_rgba, mine_depth, want_rgba, want_depth):
    """north-star tolerance terms on the float canvas: share of pixels within 1/255 on every RGBA
    channel, largest deviation in 1/255 units, PSNR (dB, peak 1.0), bit equality"""
    a = np.asarray(mine_rgba, np.float32).reshape(-1, 4)
    b = np.asarray(want_rgba, np.float32).reshape(-1, 4)
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    mse = float((d * d).mean())
    same = bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
    covered = (b[:, 3] > 0) | (a[:, 3] > 0)
    dd = np.asarray(mine_depth, np.float32).reshape(-1)[covered]
    wd = np.asarray(want_depth, np.float32).reshape(-1)[covered]
    return {"within1": float((d.max(axis=1) <= 1.0 / 255.0 + 1e-9).mean()), "max": float(d.max() * 255.0),
            "psnr": (10.0 * float(np.log10(1.0 / mse))) if mse > 0 else 999.0,
            "bit_exact": same, "bit_identical_pixels": float((a.view(np.uint32) == b.view(np.uint32)).all(axis=1).mean()),
            "depth_bit_exact_where_covered": bool(np.array_equal(dd.view(np.uint32), wd.view(np.uint32))),
            "pixels": int(a.shape[0]), "covered_pixels": int(covered.sum()),
            "ok": bool((d.max(axis=1) <= 1.0 / 255.0 + 1e-9).mean() >= 0.999 and d.max() <= 3.0 / 255.0 + 1e-9
                       and (mse == 0 or 10.0 * np.log10(1.0 / mse) >= 50.0)),
            "against": "oracle/ (CPU restatement of the reference path), same inputs, outside the timed region; "
                       "psnr 999 = identical"}


def cpu_entry(dt, n, cores, wl, what):
    return {"value": wl["W"] * wl["H"] / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "ms_per_frame": dt * 1e3,
            "sample": "%d whole frame(s)/view(s) of the same workload on the host cores (%s; oracle port of the "
                      "reference's VTK-m/OpenMP algorithm -- the reference itself is not buildable here)" % (n, what)}


# ----------------------------------------------------------------------------- CPU arm
def run_reference(args, wl):
    """The reference's CPU path (oracle restatement, all host threads) on the same workload.  Nothing of
    the product is imported here: inputs, camera, table and sample distance all come from oracle/."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as O
    cores = O.use_all_cores()
    t0 = time.time()
    fields = []
    for b in block_layout(wl):
        fields.append(O.braid_values(*b["dims"], *b["start"], *b["glob"], dtype=np.float32))
    gen_s = time.time() - t0
    sc = OracleScene(wl, fields, args.gpus)
    views = wl.get("views", 1)
    # one step = one frame; for the cinema workload one step = a bounded sample of its views
    per_step = min(views, 4)
    for _ in range(args.warmup):
        sc.frame(0)
    t0 = time.time()
    for k in range(args.steps):
        for v in range(per_step):
            sc.frame(k * per_step + v)
    dt = (time.time() - t0) / (args.steps * per_step)  # per frame (view)
    W, H = wl["W"], wl["H"]
    val = W * H / dt / 1e6
    what = "every step = %d of the %d views" % (per_step, views) if views > 1 else "every step = one whole frame"
    line = {"impl": "reference", "metric": "volume_render_mrays_per_s", "value": val, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * views * 1e3,
            "higher_is_better": True, "scaling": "strong" if wl["key"] == "c3" and under_torchrun() else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, args.gpus),
            "frames_per_s": 1.0 / dt,
            "cpu_baseline": dict(cpu_entry(dt, args.steps * per_step, cores, wl,
                                           what + "; field generation %.1fs excluded" % gen_s), value=val),
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm, one GPU
def scene_params(wl, rng_minmax):
    """Host-side driver values of VolumeRenderer::PreExecute for the workload (product side)."""
    from ascent_b200 import _lib, camera, color_table
    blocks = block_layout(wl)
    bl = [block_bounds(wl, b) for b in blocks]
    gb = union_bounds(bl)
    cam = camera.Camera().reset_to_bounds(gb)
    cams = None
    if wl.get("views"):
        cams = [c.to_struct() for c in camera.cinema_cameras(gb, *camera.cinema_angles(8, 8))]
    lut = color_table.parse_color_table(RAMP_TF).corrected_opacity(SAMPLES).lut()
    return dict(bounds=bl, gb=gb, cam=cam.to_struct(), cams=cams, lut=lut,
                sample_dist=_lib.sample_distance(gb, SAMPLES), rmin=rng_minmax[0], rmax=rng_minmax[1])


def traffic_entry(key):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        return json.load(open(tp)).get(key)
    return None


def roofline_of(alg_bytes_per_launch, kernel_ms_per_launch, launches_per_step, traffic_key, kernel):
    """achieved = algorithmic bytes per launch / average launch duration.  `frac` is that
    accounting fraction (at samples = 100 the rays skip most of the block, so it is NOT a bandwidth
    statement); `frac_measured` is the DRAM traffic ncu counted for the same launch / the same duration."""
    pk, pk_src = peaks()
    achieved = alg_bytes_per_launch / (kernel_ms_per_launch * 1e-3) / 1e9
    t = traffic_entry(traffic_key)
    traffic = t["traffic_bytes"] if t else None
    out = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": pk["hbm_gbs"], "peak_source": pk_src,
           "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
           "traffic_source": t["source"] if t else None,
           "achieved_measured": (traffic / (kernel_ms_per_launch * 1e-3) / 1e9) if traffic else None,
           "frac_measured": (traffic / (kernel_ms_per_launch * 1e-3) / 1e9 / pk["hbm_gbs"]) if traffic else None,
           "algorithmic_bytes_per_launch": alg_bytes_per_launch, "kernel_ms_per_launch": kernel_ms_per_launch,
           "launches_per_step": launches_per_step}
    if out["frac"] > 1.2 and traffic is None:
        out["note"] = "frac > 1.2 without a measured traffic figure: accounting only, not evidence"
    return out


def measure_single(args, wl, full=True):
    """One GPU, one process: the whole workload on cuda:0.  full: also e2e, cpu_baseline, parity."""
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
    sp = scene_params(wl, (rmin, rmax))
    ctx.set_tf(sp["lut"])
    axes = [warp_axis(n) for n in blocks[0]["dims"]] if wl.get("rectilinear") else None
    for i, b in enumerate(blocks):
        if axes:
            ctx.block_rectilinear(i, b["dims"], axes, None, device_ptr=fields[i].data_ptr(), dtype=_lib.VR_F32)
        else:
            ctx.block_uniform(i, b["dims"], b["origin"], b["spacing"], None, device_ptr=fields[i].data_ptr(),
                              dtype=_lib.VR_F32)
    cam = sp["cam"]
    multi = len(blocks) > 1
    views = sp["cams"] or [cam]
    ids = list(range(len(blocks)))

    def frame(ev=None, view_list=None):
        if not multi:
            # RenderOneDomainPerRank on a cleared canvas: Canvas::Clear + RenderCells + Image::Init +
            # ImageToCanvas in one launch (vr_trace_to_image); the cinema workload renders all its
            # views back to back (the reference batches <= 10 renders per Update, Scene.cpp:133-149)
            if ev:
                ev[0].record(stream)
            for v in (view_list or views):
                ctx.trace_to_image(0, v, W, H, sp["sample_dist"], rmin, rmax, write_canvas=True)
            if ev:
                ev[1].record(stream)
        else:
            # RenderMultipleDomainsPerRank on dense ray layers: the whole block loop is one ABI call,
            # then PartialCompositor::composite + partials_to_canvas over a cleared canvas in ONE kernel
            ctx.layers_begin(W, H)
            if ev:
                ev[0].record(stream)
            ctx.trace_blocks_to_layers(ids, cam, sp["sample_dist"], rmin, rmax, False)
            if ev:
                ev[1].record(stream)
            ctx.layers_composite_to_canvas(cam, canvas_is_clear=True)

    warm = max(args.warmup, 3)
    with torch.cuda.stream(stream):
        for _ in range(warm):
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
        ctx.comm_join()  # the last frame's fold runs on the library's exchange stream: inside the timed region
        t_end.record(stream)
        torch.cuda.synchronize()
        clk = clocks.stop()
        launches = ctx.kernel_launches() - l0
        total_ms = t_begin.elapsed_time(t_end)
        trace_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    ms = total_ms / args.steps
    value = W * H * len(views) / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (the sampler), per launch
    n_launch = len(blocks) * len(views)
    # SURVEY 8(d), unit of work = one frame of one block: N_vox*4 + (pixels written)*20.  One launch
    # traces one block for one view; a multi-block frame writes its canvas once, in the fold
    alg_launch = nvox * 4 + (W * H * 20 if not multi else 0)
    if multi:
        from ascent_b200 import _lib as L
        alg_launch += int(np.mean([max(0, s[2]) * max(0, s[3]) for s in
                                   (L.find_subset(cam, W, H, b) for b in sp["bounds"])])) * 20
    roof = roofline_of(alg_launch, trace_ms / n_launch if multi else trace_ms / len(views), n_launch,
                       wl["key"] + ("_p1" if SAMPLES == 887 else "") if SAMPLES in (100, 887) else "none",
                       "trace_kernel (sampler.cu)")
    if multi:
        roof["note_overlap"] = ("the %d per-block launches of a frame overlap on side streams: kernel_ms_per_launch is "
                                "the block loop's span / launches, a lower bound of a launch's own duration" % n_launch)

    line = {"metric": "volume_render_mrays_per_s", "value": value, "unit": "Mrays/s", "n_gpus": 1,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if wl["key"] == "c3" and under_torchrun() else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, 1),
            "frames_per_s": 1e3 * len(views) / ms, "render_ms_per_frame": trace_ms / len(views),
            "composite_ms_per_frame": (ms - trace_ms) / len(views),
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof}
    if not full:
        ctx.close()
        return line

    # ---- parity + CPU baseline: the oracle on the very same field values (downloaded from the GPU)
    host_fields = [f.cpu().numpy() for f in fields]
    parity = cpu = None
    if not args.no_cpu:
        from oracle import oracle as O
        cores = O.use_all_cores()
        sc = OracleScene(wl, host_fields, 1, rng=(rmin, rmax))
        check_views = [0, 9] if len(views) > 1 else [0]
        results = []
        for v in check_views:
            with torch.cuda.stream(stream):
                frame(view_list=[views[v]] if not multi else None)
                g_rgba, g_depth = ctx.canvas_download(W, H)
            o_rgba, o_depth = sc.frame(v)
            results.append(compare_canvas(g_rgba, g_depth, o_rgba, o_depth))
        parity = results[0] if len(results) == 1 else dict(
            results[0], views_checked=check_views, ok=all(r["ok"] for r in results),
            within1=min(r["within1"] for r in results), max=max(r["max"] for r in results),
            psnr=min(r["psnr"] for r in results), bit_exact=all(r["bit_exact"] for r in results))
        dt, n = sc.time_frames(20, 10.0)
        cpu = cpu_entry(dt, n, cores, wl, "bounded to ~10 s")
        if multi:
            line["partials_total"] = sc.n_partials

    # ---- e2e: host buffers through the ABI a vtk-h caller uses, H2D + D2H inside the timed region
    e2e = e2e_single(args, ctx, stream, wl, blocks, fields, sp, rmin, rmax, axes)
    line.update({"e2e": e2e, "parity": parity, "cpu_baseline": cpu})
    if wl["key"] == "c2":
        line["png"] = png_epilogue(ctx, stream, frame, W, H)
    ctx.close()
    return line


def png_epilogue(ctx, stream, frame, W, H, reps=10):
    """Render::Save's part of the frame (Render.cpp:299-312): the finished canvas as a PNG file in host memory.
    Three ways, each timed per call through the C ABI (the calls synchronise): the file encoded on the device
    (vr_canvas_encode_png: only the file crosses PCIe), the RGBA8 frame downloaded (vr_canvas_download_rgba8,
    what the reference's encoder starts from) and, for scale, zlib at level 1 on the host over those bytes --
    lodepng itself (Huffman only) is not available here."""
    import io
    import zlib
    import torch
    try:
        with torch.cuda.stream(stream):
            frame()
            ctx.synchronize()
            bg = np.array([1.0, 1.0, 1.0, 1.0], np.float32)
            from ascent_b200 import _lib as L
            file_t = torch.empty(int(L.load().vr_png_bound(W, H)), dtype=torch.uint8).pin_memory()  # (pinned, like the
            out_t = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()                         # e2e's host buffers)
            png = ctx.canvas_encode_png(W, H, bg, out=file_t.numpy())  # warm (buffers)
            t = []
            for _ in range(reps):
                t0 = time.perf_counter()
                png = ctx.canvas_encode_png(W, H, bg, out=file_t.numpy())
                t.append(time.perf_counter() - t0)
            png = png.tobytes()
            out = out_t.numpy()
            ctx.canvas_download_rgba8(W, H, bg, flip=True, out=out)
            t2 = []
            for _ in range(reps):
                t0 = time.perf_counter()
                ctx.canvas_download_rgba8(W, H, bg, flip=True, out=out)
                t2.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        z = zlib.compress(out.tobytes(), 1)
        host_ms = (time.perf_counter() - t0) * 1e3
        ok = None
        try:
            from PIL import Image
            ok = bool(np.array_equal(np.array(Image.open(io.BytesIO(png)).convert("RGBA")), out))
        except Exception:
            pass
        return {"device_encode_ms_per_call": float(np.median(t)) * 1e3, "file_bytes": len(png),
                "raw_rgba8_bytes": int(W * H * 4), "download_rgba8_ms_per_call": float(np.median(t2)) * 1e3,
                "host_zlib_level1_ms": host_ms, "host_zlib_bytes": len(z), "decodes_to_the_rgba8_frame": ok,
                "what": "vr_canvas_encode_png: background blend + float->uint8 + flip + PNG (Sub filter, per-scanline "
                        "fixed-Huffman deflate with run-length matches, Adler/CRC combined from per-row partials) on the "
                        "GPU, file copied to the host; per call incl. two stream synchronisations"}
    except Exception as e:  # (diagnostic key: never fails the bench line)
        return {"error": repr(e)}


def footprint_rect(cam, W, H, bounds_list):
    """(x0, y0, x1, y1): the screen rectangle outside of which a frame that started from Canvas::Clear IS the cleared
    canvas -- the union of vr_find_subset over all domains, x widened to the 4-pixel groups the fused kernels write"""
    from ascent_b200 import _lib as L
    x0, y0, x1, y1 = W, H, 0, 0
    for b in bounds_list:
        sx, sy, sw, sh = L.find_subset(cam, W, H, b)
        if sw <= 0 or sh <= 0:
            continue
        x0, y0 = min(x0, sx & ~3), min(y0, sy)
        x1, y1 = max(x1, min(W, (sx + sw + 3) & ~3)), max(y1, sy + sh)
    return (x0, y0, x1, y1) if x1 > x0 and y1 > y0 else (0, 0, 0, 0)


def union_rect(a, b):
    if a is None or a[2] <= a[0]:
        return b
    if b is None or b[2] <= b[0]:
        return a
    return (min(a[0], b[0]), min(a[1], b[1]), max(a[2], b[2]), max(a[3], b[3]))


def e2e_single(args, ctx, stream, wl, blocks, fields, sp, rmin, rmax, axes):
    """per step: publish every block from pinned host memory, render, read the canvas back -- all through
    the C ABI with host buffers.  copy: VR_HOST (cudaMemcpyAsync of the whole block); staged:
    VR_HOST_STAGED (sampler pre-pass flags the 128-byte lines the rays touch, a gather kernel pulls
    exactly those across PCIe; L2 flushed every step); mapped (single block only): VR_HOST_MAPPED."""
    import torch
    from ascent_b200 import _lib
    W, H = wl["W"], wl["H"]
    nvox = int(np.prod(blocks[0]["dims"]))
    multi = len(blocks) > 1
    cam = sp["cam"]
    views = sp["cams"] or [cam]
    hosts = []
    for f in fields:
        h = torch.empty(nvox, dtype=torch.float32, pin_memory=True)
        h.copy_(f)
        hosts.append(h)
    hf = [h.numpy() for h in hosts]
    rgba_h = torch.zeros(H * W * 4, dtype=torch.float32, pin_memory=True)
    depth_h = torch.full((H * W,), 1.001, dtype=torch.float32, pin_memory=True)
    hr, hd = rgba_h.numpy().reshape(-1, 4), depth_h.numpy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    base = 100

    def publish(i, mode):
        b = blocks[i]
        kw = dict(host_mapped=(mode == "mapped"), staged=(mode == "staged"))
        if axes:
            ctx.block_rectilinear(base + i, b["dims"], axes, hf[i], **kw)
        else:
            ctx.block_uniform(base + i, b["dims"], b["origin"], b["spacing"], hf[i], **kw)

    # The host canvas (hr, hd) starts as the cleared canvas -- what Render::ClearCanvas leaves -- and every frame
    # starts from Canvas::Clear on the device, so a frame differs from what the host already holds only inside its
    # footprint and the previous frame's: vr_canvas_download_rect moves those pixels only (20 B each).
    rects = [footprint_rect(v, W, H, sp["bounds"]) for v in views]
    state = {"prev": None, "d2h": 0}

    def read_back(k):
        r = union_rect(state["prev"], rects[k])
        ctx.canvas_download_rect(r, hr, hd)
        state["prev"] = rects[k]
        state["d2h"] += max(0, r[2] - r[0]) * max(0, r[3] - r[1]) * 20

    def e2e_frame(mode):
        # for the in-place modes L2 is flushed first so that no line of the previous step's
        # (identical) field can be served from cache
        if mode != "copy":
            with torch.cuda.stream(stream):
                flush.zero_()
        for i in range(len(blocks)):
            publish(i, mode)
        if not multi:
            for k, v in enumerate(views):
                ctx.trace_to_image(base, v, W, H, sp["sample_dist"], rmin, rmax, write_canvas=True)
                read_back(k)
        else:
            ctx.layers_begin(W, H)
            ctx.trace_blocks_to_layers([base + i for i in range(len(blocks))], cam, sp["sample_dist"], rmin, rmax,
                                       False)
            ctx.layers_composite_to_canvas(cam, canvas_is_clear=True)
            read_back(0)

    n_e2e = max(3, min(args.steps, 10 if not multi else 4))
    modes, moved = {}, {}
    names = ("copy", "mapped", "staged") if not multi else ("copy", "staged")
    for name in names:
        if name == "mapped" and len(views) > 8:
            continue  # many views per publish: in-place sampling re-reads the block for every view
        for _ in range(2):
            e2e_frame(name)
        state["d2h"] = 0
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_frame(name)
        modes[name] = (time.perf_counter() - t0) / n_e2e
        d2h_step = state["d2h"] // n_e2e
        if name == "staged":
            moved[name] = sum(ctx.block_staged_bytes(base + i) for i in range(len(blocks)))
    best = min(modes, key=modes.get)
    dt = modes[best]
    # the host canvas assembled from footprint read-backs IS the frame: compare with a full download
    full_r = np.empty_like(hr)
    full_d = np.empty_like(hd)
    ctx.canvas_download(W, H, full_r, full_d)
    extra = {"host_canvas_equals_full_download": bool(np.array_equal(full_r, hr) and np.array_equal(full_d, hd)),
             "d2h_full_canvas_bytes": W * H * 20 * len(views)}
    if not multi:
        # the same frame through the canvas-in/canvas-out form vtk-h's RenderCells seam needs when
        # opaque geometry is already on the canvas (upload + K2 depth clamp + blend over + download)
        publish(0, "copy")
        t0 = time.perf_counter()
        for _ in range(3):
            hr.fill(0.0)
            hd.fill(1.001)
            ctx.render_image(base, cam, W, H, sp["sample_dist"], rmin, rmax, hr, hd)
        extra["render_image_canvas_inout_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        # ... and with the whole float canvas read back every view (round 1's e2e)
        publish(0, "staged")
        t0 = time.perf_counter()
        for _ in range(3):
            with torch.cuda.stream(stream):
                flush.zero_()
            publish(0, "staged")
            for v in views:
                ctx.trace_to_image(base, v, W, H, sp["sample_dist"], rmin, rmax, write_canvas=True)
                ctx.canvas_download(W, H, hr, hd)
        extra["staged_full_canvas_readback_ms"] = (time.perf_counter() - t0) / 3 * 1e3
    t = traffic_entry(wl["key"])
    total = nvox * 4 * len(blocks)
    h2d = {"copy": total, "mapped": (t or {}).get("dram_read_bytes", total), "staged": moved.get("staged", total)}
    out = {"value": W * H * len(views) / dt / 1e6, "unit": "Mrays/s", "ms_per_step": dt * 1e3, "mode": best,
           "h2d_bytes_per_step": h2d[best], "d2h_bytes_per_step": int(d2h_step),
           "modes_ms_per_step": {k: v * 1e3 for k, v in modes.items()},
           "modes_h2d_bytes": {k: h2d[k] for k in modes}, "field_bytes": total,
           "what": "per step: vr_block_*(pinned host field) for every block + render + vr_canvas_download_rect (the "
                   "frame's screen footprint of the float canvas into the host canvas, which holds the cleared canvas "
                   "elsewhere -- checked against a full download after the run) per view, all through the C ABI with "
                   "host buffers; the fastest publish mode is the headline, images are bit-identical in all of them "
                   "(tests/test_gpu_parity.py)"}
    out.update(extra)
    for i in range(len(blocks)):
        ctx.block_free(base + i)
    return out


def run_single(args, wl):
    line = measure_single(args, wl, full=True)
    if wl["key"] == "c2" and not args.workload and not args.no_c3:
        # the scaling curve's T(1): config 3 on this one GPU (what `torchrun ... --gpus 1` runs)
        wl3 = workload_c3()
        wl3["name"] = wl3["name"].replace("samples=100", "samples=%d" % SAMPLES)
        l3 = measure_single(args, wl3, full=True)
        line["c3_n1"] = {k: l3.get(k) for k in ("value", "unit", "ms_per_step", "frames_per_s", "render_ms_per_frame",
                                                 "composite_ms_per_frame", "config", "e2e", "parity",
                                                 "cpu_baseline", "partials_total", "gpu_launches")}
    print(json.dumps(line))


def main():
    global SAMPLES
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample and the parity check")
    ap.add_argument("--no-c3", action="store_true", help="plain N=1 run: skip the extra c3_n1 measurement")
    ap.add_argument("--no-e2e", action="store_true", help="diagnostic runs: skip the end-to-end (host buffer) measurement")
    ap.add_argument("--one-block-per-rank", action="store_true",
                    help="N > 1 diagnostics: rank r renders only block r of c3 (path A at any N); not a bench line")
    ap.add_argument("--nccl-baseline", action="store_true",
                    help="N > 1, image path: also time the NCCL all_to_all + fold + gather form of the exchange")
    ap.add_argument("--samples", type=int, default=None,
                    help="samples option of the volume plot (default 100; 887 = one sample per voxel on c2)")
    args = ap.parse_args()
    if args.samples:
        SAMPLES = args.samples
    # one workload per launch style: a plain `--gpus 1` run is config 2; anything started by torchrun
    # (the 1/2/4/8 scaling sweep, N = 1 included) is config 3
    wname = args.workload or ("c3" if (args.gpus > 1 or under_torchrun()) else "c2")
    wl = WORKLOADS[wname]()
    wl["name"] = wl["name"].replace("samples=100", "samples=%d" % SAMPLES)
    if args.impl == "reference":
        return run_reference(args, wl)
    if args.gpus == 1:
        return run_single(args, wl)
    from ascent_b200 import distributed
    return distributed.run_bench(args, wl, sys.modules[__name__])


if __name__ == "__main__":
    main()
