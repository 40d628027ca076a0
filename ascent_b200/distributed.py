"""One process per GPU: the host-side plumbing of the sort-last path.

What vtk-h does with MPI around the volume renderer (global bounds / scalar range Allreduce,
Renderer.cpp:144-179; visibility ordering gather + sort + scatter, VolumeRenderer.cpp:690-867;
"one domain per rank?" Allreduce, :468-480) is done here with ``torch.distributed`` -- NCCL on the
GPU box, gloo in the CPU tests.  Only O(ranks) scalars travel this way; pixels and partials move
GPU-to-GPU inside the compositing kernels (csrc/comm.cu) through the IPC-mapped exchange arenas
whose handles are all-gathered once at start-up.
"""
import os
import time

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------- pure host logic
def assign_blocks(n_blocks, world_size, mode="round_robin"):
    """Which blocks a rank renders.  Sort-last rendering does not care (Ascent renders whatever
    domains the simulation left on each rank), load balance does: with blocks numbered x-fastest,
    "round_robin" (rank r owns r, r + world, ...) gives every rank the same number of blocks near
    the camera and far from it for any axis-aligned view, where "contiguous" (rank r owns
    [r*n/world, (r+1)*n/world), DIY's ContiguousAssigner) hands one rank all the front blocks."""
    if mode == "contiguous":
        return [list(range(r * n_blocks // world_size, (r + 1) * n_blocks // world_size))
                for r in range(world_size)]
    return [list(range(r, n_blocks, world_size)) for r in range(world_size)]


def one_domain_per_rank(n_local, dist=None):
    """VolumeRenderer::DoExecute's path switch (VolumeRenderer.cpp:468-480, OneDomainPerRank):
    path A only when EVERY rank holds exactly one domain."""
    flag = 1 if n_local == 1 else 0
    if dist is None:
        return bool(flag)
    import torch
    t = torch.tensor([flag], dtype=torch.int32, device=_dist_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.item()))


def _dist_device(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def all_gather_array(arr, dist):
    """all-gather a small fixed-shape numpy array; returns (world, *shape)."""
    import torch
    a = np.ascontiguousarray(arr)
    t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(_dist_device(dist))
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy().view(a.dtype).reshape(a.shape) for o in out])


def global_range(local_min, local_max, dist):
    """Renderer::PreExecute: global scalar range (two Allreduces in the reference)."""
    import torch
    t = torch.tensor([local_min, -local_max], dtype=torch.float64, device=_dist_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t[0].item()), float(-t[1].item())


def global_bounds(local_bounds, dist):
    """vtkh::DataSet::GetGlobalBounds: union over ranks of the union of local domain bounds."""
    import torch
    lb = np.asarray(local_bounds, np.float64).reshape(-1, 6)
    if lb.shape[0] == 0:
        v = np.array([np.inf, np.inf, np.inf, np.inf, np.inf, np.inf])
    else:
        v = np.array([lb[:, 0].min(), -lb[:, 1].max(), lb[:, 2].min(), -lb[:, 3].max(), lb[:, 4].min(),
                      -lb[:, 5].max()])
    t = torch.from_numpy(v).to(_dist_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    v = t.cpu().numpy()
    return np.array([v[0], -v[1], v[2], -v[3], v[4], -v[5]])


def global_visibility_order(local_bounds, cam, dist):
    """FindVisibilityOrdering + DepthSort (VolumeRenderer.cpp:690-867) for one camera: every rank's
    domain depths are gathered, sorted ascending (ties keep (rank, domain) order) and each rank gets
    the order index of its own domains.  The reference gathers to rank 0 and scatters; all-gather +
    the same sort on every rank gives the identical integers without the second hop.
    Every rank must hold the same number of domains (the bench and the reference's path A do)."""
    lb = np.asarray(local_bounds, np.float64).reshape(-1, 6)
    allb = all_gather_array(lb, dist).reshape(-1, 6)  # (rank, domain) order
    order = _lib.visibility_order(allb, cam)
    n = lb.shape[0]
    r = dist.get_rank()
    return order.reshape(dist.get_world_size(), n), order.reshape(dist.get_world_size(), n)[r]


def connect(ctx, dist, max_pixels, max_partials):
    """Allocate this rank's exchange arena, all-gather the CUDA IPC handles, map the peers."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    # the arena geometry must be the same on every rank: agree on the largest request
    t = torch.tensor([max_pixels, max_partials], dtype=torch.int64, device=_dist_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_pixels, max_partials = int(t[0].item()), int(t[1].item())
    h = ctx.comm_init(rank, world, max_pixels, max_partials)
    hs = all_gather_array(np.frombuffer(h, np.uint8), dist)
    dist.barrier()
    ctx.comm_connect(hs.tobytes())
    dist.barrier()


# ----------------------------------------------------------------------------- bench (N > 1)
def run_bench(args, wl, bench):
    """BASELINE config 3 on N ranks (one per GPU): 8 blocks of 512^3 over N GPUs, 3840x2160,
    sort-last render + composite to rank 0.  N == 8 -> path A (uint8 image, direct-send fused in
    one P2P kernel); N < 8 -> path B (float partials, pull + merge + fold fused in one P2P kernel),
    exactly the reference's switch."""
    import json

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "VERSION") == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    blocks = bench.block_layout(wl)
    mine = assign_blocks(len(blocks), world)[rank]
    if getattr(args, "one_block_per_rank", False):
        # diagnostics: the front blocks of c3, one per rank -> the image path (A) at any N
        blocks = blocks[len(blocks) - world:]
        mine = [rank]
        wl = dict(wl, name=wl["name"] + " [diagnostic: %d front blocks only]" % world)
    W, H = wl["W"], wl["H"]
    nvox = int(np.prod(blocks[0]["dims"]))
    fields = {}
    with torch.cuda.stream(stream):
        for i in mine:
            b = blocks[i]
            t = torch.empty(nvox, dtype=torch.float32, device="cuda")
            ctx.synth_braid_dev(t.data_ptr(), _lib.VR_F32, b["dims"], b["start"], b["glob"])
            fields[i] = t
        stream.synchronize()
        lmin = min(float(t.min()) for t in fields.values())
        lmax = max(float(t.max()) for t in fields.values())
    rmin, rmax = global_range(lmin, lmax, dist)
    sp = bench.scene_params(wl, (rmin, rmax))
    if getattr(args, "one_block_per_rank", False):
        sp["bounds"] = sp["bounds"][len(sp["bounds"]) - world:]
    ctx.set_tf(sp["lut"])
    for i in mine:
        b = blocks[i]
        ctx.block_uniform(i, b["dims"], b["origin"], b["spacing"], None, device_ptr=fields[i].data_ptr(),
                          dtype=_lib.VR_F32)
    cam = sp["cam"]
    path_a = one_domain_per_rank(len(mine), dist)
    # upper bound of partials per rank: every ray of every local block's screen subset
    max_partials = 0
    if not path_a:
        for i in mine:
            s = _lib.find_subset(cam, W, H, sp["bounds"][i])
            max_partials += s[2] * s[3] + 4  # layers start on 4-entry boundaries
    connect(ctx, dist, W * H, max_partials)
    vis_all, _ = global_visibility_order([sp["bounds"][i] for i in mine], cam, dist)
    vis_rank = np.ascontiguousarray(vis_all[:, 0], np.int32)  # path A: one domain per rank

    # path A: the sampler's epilogue stores every finished pixel straight into its owner's receive slot
    # (VR_FRAME_PUSH), so the exchange folds from local memory; VR_PUSH=0 keeps the image local (pull)
    use_push = os.environ.get("VR_PUSH", "1") == "1" and W % 4 == 0

    def render(ahead=False):
        if path_a:
            # Canvas::Clear + RenderCells + Image::Init in one launch, straight into the exchange arena
            # (ahead: into the next slot of the image ring, see VR_FRAME_AHEAD)
            ctx.trace_to_image(mine[0], cam, W, H, sp["sample_dist"], rmin, rmax, no_clear=True, ahead=ahead,
                               push=use_push)
        else:
            ctx.layers_begin(W, H)
            ctx.trace_blocks_to_layers(mine, cam, sp["sample_dist"], rmin, rmax, False)

    def composite():
        if path_a:
            ctx.comm_composite_images_to_canvas(vis_rank)
        else:
            # redistribute + sort + fold + collect + partials_to_canvas: ONE P2P kernel per rank that
            # gathers each owned pixel's entries from every rank's ray layers over NVLink, folds
            # them in depth order and stores finished pixels into rank 0's canvas
            ctx.comm_layers_composite_to_canvas(cam)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    warm = max(args.warmup, 3)
    with torch.cuda.stream(stream):
        for _ in range(warm):
            render()
            composite()
        torch.cuda.synchronize()
        dist.barrier()
        l0 = ctx.kernel_launches()
        clocks = bench.ClockSampler(local) if rank == 0 else None
        if clocks:
            clocks.start()
        def timed(pipelined):
            """K steps, each one trace + one exchange.  serial: trace(k), exchange(k).  pipelined (path A,
            the renders of a batch are independent, Scene.cpp:133-149): trace(k+1) is issued BEFORE
            exchange(k) -- the first trace is the prologue outside the timed region, the last traced
            image is left over -- so a rank that finishes early traces on instead of idling in the
            exchange.  Same kernels and the same number of them per step either way."""
            marks = [(ev(), ev(), ev()) for _ in range(args.steps)]
            t0, t1 = ev(), ev()
            if pipelined:
                render()
            torch.cuda.synchronize()
            dist.barrier()
            h0 = time.perf_counter()
            t0.record(stream)
            for k in range(args.steps):
                marks[k][0].record(stream)
                render(ahead=pipelined)
                marks[k][1].record(stream)
                composite()
                marks[k][2].record(stream)
            ctx.comm_join()  # the last exchange runs on the library's exchange stream: inside the timed region
            t1.record(stream)
            host_issue[0] = (time.perf_counter() - h0) * 1e3 / args.steps
            torch.cuda.synchronize()
            dist.barrier()
            return (t0.elapsed_time(t1), float(np.mean([a.elapsed_time(b) for a, b, _ in marks])),
                    float(np.mean([b.elapsed_time(c) for _, b, c in marks])))

        def timed_batch():
            """the same K frames through vr_comm_render_frames: ONE ABI call issues every trace + exchange of the
            batch from C++ (what a vtk-h caller's loop over the renders of a batch does), no per-frame Python"""
            cams = (_lib.CameraStruct * args.steps)(*[_lib.as_camera(cam) for _ in range(args.steps)])
            vo = np.tile(vis_rank, (args.steps, 1))
            t0, t1 = ev(), ev()
            torch.cuda.synchronize()
            dist.barrier()
            h0 = time.perf_counter()
            t0.record(stream)
            ctx.comm_render_frames(mine[0], cams, W, H, sp["sample_dist"], rmin, rmax, vo)
            ctx.comm_join()
            t1.record(stream)
            host_issue[1] = (time.perf_counter() - h0) * 1e3 / args.steps
            torch.cuda.synchronize()
            dist.barrier()
            return t0.elapsed_time(t1)

        host_issue = [0.0, 0.0]
        # NVLink payload bytes of this GPU from the NVML hardware counters, around the K timed frames
        dev_index = torch.cuda.current_device()
        nv0 = bench.nvlink_counters(dev_index)
        serial = timed(False)
        nv1 = bench.nvlink_counters(dev_index)
        nv_meas = [-1.0, -1.0]
        if nv0 is not None and nv1 is not None:
            nv_meas = [(nv1[0] - nv0[0]) / args.steps, (nv1[1] - nv0[1]) / args.steps]
        host_serial = host_issue[0]
        launches = ctx.kernel_launches() - l0
        piped = timed(True) if path_a else None
        batch_ms = None
        if path_a and use_push and os.environ.get("VR_BATCH", "1") == "1":
            timed_batch()  # (warm: camera array set-up, first use of the entry point)
            bt = torch.tensor([timed_batch()], dtype=torch.float64, device="cuda")
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
            batch_ms = float(bt[0])
        clk = clocks.stop() if clocks else None
        # the pipelined order is the product path for batches of renders; report it when it wins on
        # EVERY rank's clock (max over ranks is taken below), the serial order otherwise
        use_piped = False
        if piped is not None:
            tp = torch.tensor([piped[0], serial[0]], dtype=torch.float64, device="cuda")
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            use_piped = bool(tp[0] < tp[1])
            serial_total_max = float(tp[1])
            piped_total_max = float(tp[0])
        total_ms, render_ms, tail_ms = piped if use_piped else serial
        order_used = "pipelined" if use_piped else "serial"
        if batch_ms is not None:
            tb = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            if batch_ms < float(tb[0]):
                total_ms, order_used = batch_ms, "batch"

        # trace alone and composite alone: ranks aligned by a barrier, nothing else in flight (the library
        # runs image-only traces and the exchange on its own streams: comm_join brings them back)
        comp, rend = [], []
        for _ in range(min(args.steps, 10)):
            torch.cuda.synchronize()
            dist.barrier()
            a, b = ev(), ev()
            a.record(stream)
            render()
            ctx.comm_join()
            b.record(stream)
            torch.cuda.synchronize()
            rend.append(a.elapsed_time(b))
            dist.barrier()
            a, b = ev(), ev()
            a.record(stream)
            composite()
            ctx.comm_join()
            b.record(stream)
            torch.cuda.synchronize()
            comp.append(a.elapsed_time(b))
        comp_ms = float(np.median(comp))
        render_ms = float(np.median(rend))
        n_partials = 0 if path_a else _count_local_partials(ctx, render)

    t = torch.tensor([total_ms, render_ms, tail_ms, comp_ms, float(launches), float(n_partials), nv_meas[0], nv_meas[1]],
                     dtype=torch.float64, device="cuda")
    per_rank = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(per_rank, t)
    nv_rows = [[float(pr[6]), float(pr[7])] for pr in per_rank]
    per_rank = [[round(float(x), 4) for x in pr[:4]] for pr in per_rank]
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    total_ms, render_ms, tail_ms, comp_ms = [float(x) for x in tmax[:4]]
    ms = total_ms / args.steps

    # ---- diagnostics (VR_TIMELINE=1): where the time of one exchange goes on every rank, from the
    # globaltimer stamps the kernels leave in the arena flags (per GPU; differences only)
    timeline = None
    if path_a and os.environ.get("VR_TIMELINE") == "1":
        rows = []
        with torch.cuda.stream(stream):
            for _ in range(8):
                torch.cuda.synchronize()
                dist.barrier()
                render()
                composite()
                tl = ctx.comm_timeline()
                rows.append([(tl[0] - tl[6]) * 1e-3, (tl[1] - tl[0]) * 1e-3, (tl[2] - tl[1]) * 1e-3,
                             (tl[3] - tl[2]) * 1e-3 if rank == 0 else 0.0, (tl[4] - tl[3]) * 1e-3 if rank == 0 else 0.0,
                             (tl[5] - tl[4]) * 1e-3 if rank == 0 else 0.0,
                             (tl[8] - tl[1]) * 1e-3, (tl[9] - tl[8]) * 1e-3, (tl[10] - tl[9]) * 1e-3,
                             (tl[11] - tl[10]) * 1e-3, (tl[2] - tl[11]) * 1e-3])
        tt = torch.tensor(np.median(np.array(rows[2:]), axis=0), dtype=torch.float64, device="cuda")
        allt = [torch.empty_like(tt) for _ in range(world)]
        dist.all_gather(allt, tt)
        timeline = {"columns": ["trace_end_to_fold_start", "wait_all_ready", "fold", "fold_end_to_canvas_start(rank0)",
                                "wait_all_done(rank0)", "to_canvas(rank0)", "cta0:prologue", "cta0:own_chunks",
                                "cta0:rank0_clears", "cta0:fence+count", "cta0_end_to_last_cta_out"],
                    "rows_us": [[round(float(x), 2) for x in r] for r in allt]}

    nccl_base = None
    if path_a and getattr(args, "nccl_baseline", False):
        def render_full():
            ctx.trace_to_image(mine[0], cam, W, H, sp["sample_dist"], rmin, rmax)
        nccl_base = nccl_image_composite_baseline(ctx, dist, stream, W, H, vis_rank, render_full)

    # ---- e2e: every rank publishes its blocks from pinned host memory, root reads the canvas back
    e2e = None
    if not getattr(args, "no_e2e", False):
        e2e = _e2e(ctx, dist, stream, blocks, mine, fields, sp, cam, W, H, rmin, rmax, render, composite,
                   rank, max(3, min(args.steps, 5)))

    # ---- rank 0, outside every timed region: the oracle on the same inputs (parity of the frame that was
    # timed, CPU baseline on the host cores) and T(1) of the same workload on this rank's GPU alone
    parity = cpu = t1 = None
    diag = getattr(args, "one_block_per_rank", False)
    if rank == 0 and not args.no_cpu and not diag:
        with torch.cuda.stream(stream):
            render()
            composite()
            g_rgba, g_depth = ctx.canvas_download(W, H)
    else:
        with torch.cuda.stream(stream):
            if not args.no_cpu and not diag:
                render()
                composite()
            ctx.synchronize()
    if rank == 0 and not args.no_cpu and not diag:
        from oracle import oracle as O
        cores = O.use_all_cores()
        host_fields = []
        tmp = torch.empty(nvox, dtype=torch.float32, device="cuda")
        for b in blocks:
            ctx.synth_braid_dev(tmp.data_ptr(), _lib.VR_F32, b["dims"], b["start"], b["glob"])
            ctx.synchronize()
            host_fields.append(tmp.cpu().numpy())
        del tmp
        sc = bench.OracleScene(wl, host_fields, world, rng=(rmin, rmax))
        o_rgba, o_depth = sc.frame(0)
        parity = bench.compare_canvas(g_rgba, g_depth, o_rgba, o_depth)
        dt_cpu, n_cpu = sc.time_frames(5, 10.0)
        cpu = bench.cpu_entry(dt_cpu, n_cpu, cores, wl, "bounded to ~10 s, rank 0's host cores")
        del sc, host_fields
        l1 = bench.measure_single(args, dict(wl), full=False)
        t1 = {"ms_per_step": l1["ms_per_step"], "value": l1["value"], "unit": "Mrays/s",
              "render_ms_per_frame": l1["render_ms_per_frame"], "composite_ms_per_frame": l1["composite_ms_per_frame"],
              "what": "the same workload (all %d blocks, path B) on rank 0's GPU alone, measured in this run while "
                      "the other ranks wait: T(1) of the strong-scaling curve" % len(blocks)}

    if rank == 0:
        alg = nvox * 4 + W * H * 20  # SURVEY 8(d), one launch = one block, one view
        # NVLink bytes that reach the busiest GPU (rank 0) per frame: what it pulls for the pixels it
        # owns plus what the other owners store into its result
        rects = []
        for b in sp["bounds"]:
            sx, sy, sw, sh = _lib.find_subset(cam, W, H, b)
            rects.append((sx & ~3, sy, min(W, (sx + sw + 3) & ~3), sy + sh))
        cover = np.zeros((H, W), np.uint8)
        layer_px = 0
        for (x0, y0, x1, y1) in rects:
            cover[y0:y1, x0:x1] = 1
            layer_px += max(0, x1 - x0) * max(0, y1 - y0)
        covered = int(cover.sum())
        if path_a:
            pulled = layer_px / world * (world - 1) / world * 8.0   # RGBA8 + depth of the covering layers
            pushed = covered * (world - 1) / world * 8.0            # folded pixels stored into rank 0
        else:
            alg = nvox * 4 + layer_px / len(rects) * 20  # a layer launch writes its rectangle, not the frame
            pulled = layer_px / world * (world - 1) / world * 20.0  # layer entries (rgba + depth)
            pushed = covered * (world - 1) / world * 20.0           # finished canvas pixels into rank 0
        nv_bytes = pulled + pushed
        roof = bench.roofline_of(alg, render_ms / len(mine), len(mine),
                                 ("c3_a" if path_a else "c3") if bench.SAMPLES == 100 else "none",
                                 "trace_kernel (sampler.cu)")
        roof["note"] = ("per GPU; kernel_ms_per_launch = the slowest rank's render span / its launches; at samples = 100 "
                        "the rays touch a fraction of the block, so `frac` (SURVEY 8(d) accounting) may exceed 1 -- "
                        "`frac_measured` is the DRAM-traffic statement")
        cfg = bench.config_of(wl, world)
        line = {"metric": "volume_render_mrays_per_s", "value": W * H / (ms * 1e-3) / 1e6, "unit": "Mrays/s",
                "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg,
                "details": {"exchange": ("P2P direct-send fold (uint8 images, %s)" % (
                                             "pushed by the sampler into the owners' receive slots" if use_push
                                             else "pulled from the peers' arenas")) if path_a else
                                        "P2P gather + fold of dense ray layers (float partials)",
                            "blocks_per_gpu": len(mine),
                            "order": {"pipelined": "pipelined: trace(k+1) issued before exchange(k) (VR_FRAME_AHEAD)",
                                      "serial": "serial: trace(k), exchange(k), one ABI call each",
                                      "batch": "batch: the K frames through ONE vr_comm_render_frames call (trace(k), "
                                               "exchange(k) issued from C++)"}[order_used]},
                "frames_per_s": 1e3 / ms, "render_ms_per_frame": render_ms,
                "ms_per_step_serial_order": (serial_total_max / args.steps) if piped is not None else ms,
                "ms_per_step_pipelined_order": (piped_total_max / args.steps) if piped is not None else None,
                "ms_per_step_batch_call": (batch_ms / args.steps) if batch_ms is not None else None,
                "host_issue_ms_per_frame": {"serial_python_loop": host_serial, "batch_call": host_issue[1] or None},
                "composite_ms_per_frame": comp_ms, "composite_in_step_ms": tail_ms,
                "partials_total": int(tsum[5]),
                "per_rank_ms": {"columns": ["total", "render_alone", "composite_in_step", "composite_aligned"],
                                "rows": per_rank},
                "nvlink": {"bytes_into_rank0_per_frame": nv_bytes, "pulled": pulled, "pushed_into_rank0": pushed,
                           "achieved_gbs": nv_bytes / (ms * 1e-3) / 1e9, "peak_gbs": 770.0,
                           "peak_source": "measured peer copy per direction (B200_PROFILING.md)",
                           "how": "bytes from the frame's geometry (screen rectangles); time = the whole frame (ms_per_step): "
                                  "pushed pixels / layer entries cross NVLink while the trace is still running, so the "
                                  "exchange kernel's own duration no longer contains the transfer",
                           # hardware counters (NVML NVLINK_THROUGHPUT_DATA_RX/TX, payload, KiB granularity) read on
                           # every rank around the K timed frames of the serial order: bytes per frame per GPU
                           "measured": None if nv_rows[0][0] < 0 else {
                               "rx_bytes_per_frame_by_rank": [round(r[0]) for r in nv_rows],
                               "tx_bytes_per_frame_by_rank": [round(r[1]) for r in nv_rows],
                               "rank0_rx_gbs": nv_rows[0][0] / (ms * 1e-3) / 1e9,
                               "source": "NVML field values 139/138 summed over links, delta over the timed frames"}},
                "nccl_baseline": nccl_base, "exchange_timeline": timeline,
                "t1_same_run": t1,
                "e2e": e2e, "gpu_launches": int(tsum[4]), "clocks": clk,
                "roofline": roof, "parity": parity, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


class _DevBuf:
    """A raw device pointer as something torch.as_tensor can adopt (zero copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def nccl_image_composite_baseline(ctx, dist, stream, W, H, vis_rank, render_full, n_iter=10):
    """What the reference's DirectSendCompositor does (DirectSendCompositor.cpp:121-181), with NCCL as the
    transport instead of DIY/MPI: every rank cuts its full RGBA8 + depth image into `world` bands and
    all-to-alls them (ncclSend/ncclRecv groups), folds the `world` bands it received in visibility
    order -- with the SAME fold kernel the one-GPU path uses (vr_fold_images_dev) -- then the folded
    bands are gathered on rank 0 and converted to the float canvas (vr_image_to_canvas_dev).  This is
    the baseline the fused peer-memory kernel (fold_p2p_kernel) is measured against; not a product
    path.  Returns {"composite_ms": median, "matches_fused": bool | None}."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    n = W * H
    assert n % (4 * world) == 0, "baseline needs W*H divisible by 4*world"
    band = n // world
    times = []
    with torch.cuda.stream(stream):
        recv_c = torch.empty(n, dtype=torch.int32, device="cuda")
        recv_d = torch.empty(n, dtype=torch.float32, device="cuda")
        out_c = torch.empty(band, dtype=torch.int32, device="cuda")
        out_d = torch.empty(band, dtype=torch.float32, device="cuda")
        if rank == 0:
            gath_c = [torch.empty(band, dtype=torch.int32, device="cuda") for _ in range(world)]
            gath_d = [torch.empty(band, dtype=torch.float32, device="cuda") for _ in range(world)]
            res_c = torch.empty(n, dtype=torch.int32, device="cuda")
            res_d = torch.empty(n, dtype=torch.float32, device="cuda")
        for it in range(n_iter + 2):
            render_full()                       # Canvas::Clear + trace + Image::Init: a FULL image
            rgba_ptr, depth_ptr = ctx.image_ptrs()
            img_c = torch.as_tensor(_DevBuf(rgba_ptr, n, "<i4"), device="cuda")
            img_d = torch.as_tensor(_DevBuf(depth_ptr, n, "<f4"), device="cuda")
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            dist.all_to_all_single(recv_c, img_c)
            dist.all_to_all_single(recv_d, img_d)
            ctx.fold_images_dev(recv_c.data_ptr(), recv_d.data_ptr(), band, vis_rank, band, out_c.data_ptr(),
                                out_d.data_ptr())
            dist.gather(out_c, gath_c if rank == 0 else None, dst=0)
            dist.gather(out_d, gath_d if rank == 0 else None, dst=0)
            if rank == 0:
                torch.cat(gath_c, out=res_c)
                torch.cat(gath_d, out=res_d)
                ctx.image_to_canvas_dev(res_c.data_ptr(), res_d.data_ptr())
            b.record(stream)
            torch.cuda.synchronize()
            if it >= 2:
                times.append(a.elapsed_time(b))
        matches = None
        if rank == 0:
            base_rgba, base_depth = ctx.canvas_download(W, H)
        # the same frame through the fused peer-memory exchange
        render_full()
        ctx.comm_composite_images_to_canvas(vis_rank)
        if rank == 0:
            fused_rgba, fused_depth = ctx.canvas_download(W, H)
            matches = bool(np.array_equal(base_rgba, fused_rgba) and np.array_equal(base_depth, fused_depth))
        else:
            ctx.synchronize()
        dist.barrier()
    t = torch.tensor([float(np.median(times))], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"composite_ms": float(t.item()), "matches_fused": matches,
            "what": "NCCL all_to_all of image bands + local ordered fold (vr_fold_images_dev) + NCCL gather to "
                    "rank 0 + ImageToCanvas; full images, ranks aligned by a barrier (compare composite_ms_per_frame)"}


def _count_local_partials(ctx, render):
    render()
    ctx.layers_to_partials()
    return ctx.partials_count()


def _e2e(ctx, dist, stream, blocks, mine, fields, sp, cam, W, H, rmin, rmax, render, composite, rank, n):
    import torch
    nvox = int(np.prod(blocks[0]["dims"]))
    host = {}
    for i in mine:
        h = torch.empty(nvox, dtype=torch.float32, pin_memory=True)
        h.copy_(fields[i])
        host[i] = h.numpy()
    # rank 0's host canvas: the cleared canvas (Render::ClearCanvas); every frame replaces its footprint only
    rgba_h = torch.zeros(H * W * 4, dtype=torch.float32, pin_memory=True).numpy().reshape(-1, 4) if rank == 0 else None
    depth_h = torch.full((H * W,), 1.001, dtype=torch.float32, pin_memory=True).numpy() if rank == 0 else None
    import bench as bench_mod
    rect = bench_mod.footprint_rect(cam, W, H, sp["bounds"])

    def frame(staged):
        for i in mine:
            b = blocks[i]
            # publish: dense copy (VR_HOST) or demand staging (VR_HOST_STAGED: each trace first pulls
            # the 128-byte lines its rays touch across PCIe)
            ctx.block_uniform(i, b["dims"], b["origin"], b["spacing"], host[i], staged=staged)
        render()
        composite()
        if rank == 0:
            ctx.canvas_download_rect(rect, rgba_h, depth_h)
        else:
            ctx.synchronize()

    modes, moved = {}, {}
    with torch.cuda.stream(stream):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for name, staged in (("copy", False), ("staged", True)):
            frame(staged)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                if staged:
                    flush.zero_()  # no line of the previous step's identical field may come from L2
                frame(staged)
            torch.cuda.synchronize()
            dist.barrier()
            dt = (time.perf_counter() - t0) / n
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            modes[name] = float(t.item())
            if staged:
                m = torch.tensor([float(sum(ctx.block_staged_bytes(i) for i in mine))], dtype=torch.float64,
                                 device="cuda")
                dist.all_reduce(m, op=dist.ReduceOp.SUM)
                moved[name] = int(m.item())
    best = min(modes, key=modes.get)
    dt = modes[best]
    same = None
    if rank == 0:
        full_r, full_d = ctx.canvas_download(W, H)
        same = bool(np.array_equal(full_r, rgba_h) and np.array_equal(full_d, depth_h))
    # restore the zero-copy device blocks
    for i in mine:
        b = blocks[i]
        ctx.block_uniform(i, b["dims"], b["origin"], b["spacing"], None, device_ptr=fields[i].data_ptr(),
                          dtype=_lib.VR_F32)
    h2d = {"copy": nvox * 4 * len(blocks), "staged": moved.get("staged", nvox * 4 * len(blocks))}
    return {"value": W * H / dt / 1e6, "unit": "Mrays/s", "ms_per_step": dt * 1e3, "mode": best,
            "h2d_bytes_per_step": h2d[best],
            "d2h_bytes_per_step": max(0, rect[2] - rect[0]) * max(0, rect[3] - rect[1]) * 20,
            "d2h_full_canvas_bytes": W * H * 20, "host_canvas_equals_full_download": same,
            "modes_ms_per_step": {k: v * 1e3 for k, v in modes.items()}, "modes_h2d_bytes": h2d,
            "what": "every rank: vr_block_uniform(pinned host field) per local block (copy: VR_HOST dense upload; "
                    "staged: VR_HOST_STAGED, only the 128-byte lines the rays touch cross PCIe), render, P2P "
                    "composite; rank 0: vr_canvas_download_rect (the frame's screen footprint into a host canvas that holds the "
                    "cleared canvas elsewhere; checked against a full download).  The faster mode is reported"}
