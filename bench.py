#!/usr/bin/env python
"""Benchmark of the ToC3D image-backbone hot path: 6-cam samples/s on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config toc3d_fast] [--batch B]
    python bench.py --impl reference ...      # the oracle port of the reference on the host cores

One step = one backbone forward over `batch` synthetic 6-view samples per GPU (weights: seeded
random init of the EVA-ViT-L + ToC3D architecture; there are no checkpoints offline).  N>1 is weak
scaling: every rank runs its own samples and the ranks all-gather `last_feat` over NCCL (the one
collective north_star names).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from toc3d_b200.configs import CONFIGS  # noqa: E402
from toc3d_b200.synthetic import make_gumbel, make_inputs, randomize_state_dict  # noqa: E402

METRIC = "6-cam samples/sec, EVA-ViT-L+ToC3D backbone fwd"
UNIT = "samples/s"
VIEWS = 6
IMG_NORM = dict(mean=[103.530, 116.280, 123.675], std=[57.375, 57.120, 58.395], to_rgb=False)   # ToC3D_fast.py:13-14


def load_ncu_traffic(workload):
    """DRAM bytes per GEMM launch from the committed ncu capture of this workload (profiles/ncu_traffic.json,
    written by tools/ncu_traffic.py from an `ncu --metrics dram__bytes_*` pass), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p)).get(workload)
    if not d:
        return None, None
    return d["gemm_dram_bytes_per_launch"], d["source"]


def load_peaks():
    """Roofline denominators: the driver-written MEASURED_PEAKS.json when present (any reasonable key spelling),
    else the fallback of /opt/skills/guides/B200_PROFILING.md (6.65 TB/s, 1.59 PF burst / 1.4 PF sustained)."""
    fb = dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
    except Exception:
        return fb
    flat = {}

    def walk(o, pre=""):
        if isinstance(o, dict):
            for k, v in o.items():
                walk(v, pre + str(k).lower() + ".")
        elif isinstance(o, (int, float)) and not isinstance(o, bool):
            flat[pre[:-1]] = float(o)
    walk(d)

    def pick(words, avoid=()):
        for k, v in flat.items():
            if all(w in k for w in words) and not any(a in k for a in avoid) and v > 0:
                return v
        return None
    hbm = pick(("hbm",)) or pick(("copy",)) or pick(("gb",))
    sust = pick(("sustain",))
    burst = pick(("tflop",), avoid=("sustain",)) or pick(("bf16",), avoid=("sustain",))
    if hbm and hbm < 100:          # TB/s -> GB/s
        hbm *= 1000.0
    out = dict(hbm=hbm or fb["hbm"], tf_burst=burst or fb["tf_burst"], tf_sust=sust or burst or fb["tf_sust"],
               src="measured" if (hbm and (sust or burst)) else "fallback/partly measured")
    for k in ("tf_burst", "tf_sust"):
        if out[k] > 10000:         # GFLOP/s -> TFLOP/s
            out[k] /= 1000.0
    return out


# ------------------------------------------------------------------------------- algorithmic work
def algorithmic_work(cfg, kind, hw, views):
    """FLOPs of the reference semantics incl. padded window slots, MAC = 2 FLOP, and
    the prune/gather bytes (bf16-activation convention of the survey)."""
    C, Hd = cfg["embed_dim"], int(cfg["embed_dim"] * cfg["mlp_ratio"])
    H, W = hw[0] // 16, hw[1] // 16
    V, N = views, H * W
    lin = attn = gather_b = 0.0
    lin += V * N * 2 * 768 * C
    stage = -1
    for i in range(cfg["depth"]):
        g = i in cfg["global_attn_indexes"]
        ws = cfg["global_window_size"] if g else cfg["window_size"]
        n = ws * ws
        nW = V * -(-H // ws) * -(-W // ws)
        if kind == "ToC3DEVAViT" and i in cfg["pruning_loc"]:
            stage += 1
            lin += V * N * (2 * C * 256 + 2 * 256 * 64 + 2 * 64 * 2)
        if kind == "ToC3DEVAViT" and i >= cfg["pruning_loc"][0]:
            k = int(n * cfg["token_ratio"][stage])
            rows = nW * (k + 1)
            lin += rows * (8 * C * C + 6 * C * Hd)
            attn += nW * 4 * (k + 1) ** 2 * C
            gather_b += 2 * (nW * n + rows) * C * 2
        else:
            lin += nW * n * 8 * C * C + V * N * 6 * C * Hd
            attn += nW * 4 * n * n * C
    return dict(flops=lin + attn, linear=lin, attention=attn, gather_bytes=gather_b)


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        self.rows, self.proc, self.index, self.enabled = [], None, index, enabled

    def __enter__(self):
        if not self.enabled:      # only rank 0 polls NVML: eight pollers contend on the driver lock and slow every copy
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        busy = [s for s in sm if s > 0]
        return dict(sm_mhz=statistics.median(busy or sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------- reference arm (CPU)
NECK_CFG = dict(in_channels=[1024], out_channels=256, num_outs=2)      # img_neck of ToC3D_fast.py:70-74


def workload_config(name, batch, world):
    """The `config` object of the JSON line - identical in both arms (the driver compares them)."""
    kind, cfg, hw = CONFIGS[name]
    return {"workload": name, "backbone": kind, "neck": "CPFPN 1024 -> 256, 2 levels (the multi-scale feature list)",
            "views": VIEWS, "image_hw": list(hw), "batch_per_gpu": batch,
            "prev_exists": True, "weights": "random-init EVA-ViT-L + ToC3D selectors (seed 0)",
            "parallelism": "dp%d (views x batch sharded, all-gather of the feature list)" % world}


def reference_step_fn(name, views, device="cpu", autocast=False, batch=1):
    """One forward of the reference on `batch` samples of `views` views.  Prefers the reference's OWN modules
    (unmodified sources from /root/reference or the staged baseline/_ref/, imported by tests/golden/ref_import.py with
    its third-party stubs; the stable-sort / injected-noise pins do not change the work done); falls back to the
    oracle port when the sources are not present.  bench.py is one of the two places allowed to execute oracle/.
    -> (step, kind) with kind in {"reference", "port"}."""
    import contextlib
    import io
    kind, cfg, hw = CONFIGS[name]
    inp = make_inputs(batch, views, hw, seed=0)
    gn = make_gumbel(batch * views, (hw[0] // 16) * (hw[1] // 16))
    try:
        from tests.golden.ref_import import load_reference, reference_available
        have_ref = reference_available()
    except Exception:
        have_ref = False
    if have_ref:
        ns = load_reference()
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            m = (ns.ToC3DEVAViT if kind == "ToC3DEVAViT" else ns.EVA_ViT)(**cfg).eval()
        m.load_state_dict(randomize_state_dict(m.state_dict(), seed=0, bias_std=0.02))
        m = m.to(device)
        with contextlib.redirect_stdout(io.StringIO()):
            nk = ns.CPFPN(**NECK_CFG).eval()                # the reference's own img_neck (necks/cp_fpn.py), same step as the native arm
        nk.load_state_dict(randomize_state_dict(nk.state_dict(), seed=0, bias_std=0.02))
        nk = nk.to(device)
        d = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in inp.items()}
        gnd = [g.to(device) for g in gn]

        def step():
            ns.set_gumbel(gnd)
            with torch.no_grad(), torch.autocast(device_type=torch.device(device).type, dtype=torch.bfloat16, enabled=autocast):
                feats = m(d["x"]) if kind == "EVA_ViT" else m(**d).img_feats        # petr3d.py:145-179
                return nk(list(feats.values()))                                       # petr3d.py:188-190
        return step, "reference"
    if device != "cpu":
        raise RuntimeError("the oracle port is a CPU checker; the GPU-eager reference leg needs baseline/_ref (tools/stage_ref.sh)")
    from oracle import toc3d_oracle as O
    from toc3d_b200 import CPFPN, EVA_ViT, ToC3DEVAViT
    torch.manual_seed(0)
    model = (ToC3DEVAViT if kind == "ToC3DEVAViT" else EVA_ViT)(**cfg)
    sd = randomize_state_dict(model.state_dict(), seed=0, bias_std=0.02)
    sdn = randomize_state_dict(CPFPN(**NECK_CFG).state_dict(), seed=0, bias_std=0.02)
    del model

    def step():
        with torch.no_grad():
            if kind == "EVA_ViT":
                lf = O.forward_dense(sd, cfg, inp["x"])["last_feat"]
            else:
                lf = O.forward_toc3d(sd, cfg, inp["x"], inp["temp_queries"], inp["temp_ref_points"], inp["temp_vel"],
                                     inp["temp_timestamp"], inp["temp_ego_pose"], inp["ego_pose_inv"], True, gn)["last_feat"]
            return O.neck_cpfpn(sdn, lf)
    return step, "port"


def reference_views(name):
    """Views per CPU step: the whole 6-view sample at 800x320 (a few seconds per step); one view at 1600x800, where a
    6-view step takes the best part of a minute on the host cores - reported as extrapolated."""
    return VIEWS if CONFIGS[name][2][0] * CONFIGS[name][2][1] <= 320 * 800 else 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    views = reference_views(args.config)
    step, rkind = reference_step_fn(args.config, views)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    scale = VIEWS // views
    value = 1.0 / (dt * scale)
    what = {"reference": "the reference's own ToC3DEVAViT / EVA_ViT + CPFPN modules (unmodified sources, third-party imports stubbed)",
            "port": "fp32 oracle port of the reference forward + neck (reference sources not present)"}[rkind]
    sample = "%d of %d views per step, fp32, torch.no_grad, %d torch threads: %s%s" % (
        views, VIEWS, cores, what, "" if scale == 1 else "; scaled x%d (extrapolated)" % scale)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3 * scale, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, 1, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": rkind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if scale != 1:
        line["extrapolated"] = True
    emit(line)


# ------------------------------------------------------------------------------- native arm (B200)
def build_native(name, dev, view_groups=None):
    """Backbone + neck plugins with seeded random-init weights; the neck runs inside the backbone's CUDA graph."""
    from toc3d_b200 import CPFPN, EVA_ViT, ToC3DEVAViT
    kind, cfg, hw = CONFIGS[name]
    torch.manual_seed(0)
    model = (ToC3DEVAViT if kind == "ToC3DEVAViT" else EVA_ViT)(**cfg)
    model.load_state_dict(randomize_state_dict(model.state_dict(), seed=0, bias_std=0.02))
    model = model.eval().to(dev)
    model.set_image_preprocess(**IMG_NORM)       # row f3: lets forward() take the uint8 camera crops as well
    if view_groups is not None and hasattr(model, "view_groups"):
        model.view_groups = view_groups
    neck = CPFPN(**NECK_CFG)
    neck.load_state_dict(randomize_state_dict(neck.state_dict(), seed=0, bias_std=0.02))
    neck = neck.eval().to(dev)
    model.fuse_neck(neck)
    return model, neck


def run_native(args):
    import torch.distributed as dist
    from toc3d_b200 import lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.strong:
        return run_strong(args, world, rank, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2
    B = args.batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the launch stream; L2 flushed between steps; max over ranks.  N > 1:
        the all-gather of the last step (still running on its own stream) is joined and timed as a tail."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        tail = torch.cuda.Event(enable_timing=True)
        barrier()
        for s, e in ev:
            flush.zero_()
            s.record()
            fn()
            e.record()
        if gather_stream is not None:
            torch.cuda.current_stream().wait_stream(gather_stream)
        tail.record()
        barrier()
        ms = [s.elapsed_time(e) for s, e in ev]
        if gather_stream is not None:
            ms.append(ev[-1][1].elapsed_time(tail))
        tot = torch.tensor([sum(ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return tot.item(), ms

    gather_stream = torch.cuda.Stream(device=dev) if world > 1 else None

    def make_forward(model, neck, V, H, W):
        """One step through the plugin boundary: img_backbone forward, img_neck on its feature dict (petr3d.py:159-190),
        and for N > 1 the all-gather of the multi-scale feature list.  Level 1 of the CPFPN is the stride-2 subsample of
        level 0 (cp_fpn.py:190-191), so gathering level 0 (NHWC, 256 channels) gathers the whole list.  The gather of
        step i is issued on its own stream behind forward(i) and overlaps forward(i + 1) (double-buffered destination);
        `timed` joins that stream inside the timed region, so every gather is paid for."""
        gbuf = [torch.empty(world * V, H, W, NECK_CFG["out_channels"], device=dev) for _ in range(2)] if world > 1 else None
        state = {"i": 0}

        def forward(d):
            out = model(**d)
            feats = out if isinstance(out, dict) else out.img_feats
            levels = neck(list(feats.values()))
            lv0 = levels[0].permute(0, 2, 3, 1)              # contiguous NHWC rows of level 0
            if world > 1:
                if args.serial_gather:
                    dist.all_gather_into_tensor(gbuf[0], lv0)
                else:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream())
                    with torch.cuda.stream(gather_stream):
                        gather_stream.wait_event(ev)
                        dist.all_gather_into_tensor(gbuf[state["i"] % 2], lv0)
                        lv0.record_stream(gather_stream)
                    state["i"] += 1
            return lv0
        return forward

    def quick_config(name, steps):
        """Device-resident throughput of another shipped config (same step definition), for the driver's record."""
        kind, cfg, hw = CONFIGS[name]
        m, nk = build_native(name, dev)
        H, W = hw[0] // 16, hw[1] // 16
        d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in make_inputs(B, VIEWS, hw, seed=rank).items()}
        fwd = make_forward(m, nk, B * VIEWS, H, W)
        for _ in range(3):
            fwd(d)
        t, _ = timed(lambda: fwd(d), steps)
        work = algorithmic_work(cfg, kind, hw, B * VIEWS)
        r = {"value": world * B * steps / (t * 1e-3), "unit": UNIT, "ms_per_step": t / steps, "steps": steps, "image_hw": list(hw),
             "whole_step_tflops_per_gpu": work["flops"] / (t / steps * 1e-3) / 1e12}
        del m, nk, fwd, d
        torch.cuda.empty_cache()
        return r

    kind, cfg, hw = CONFIGS[args.config]
    model, neck = build_native(args.config, dev, args.view_groups)
    V = B * VIEWS
    H, W = hw[0] // 16, hw[1] // 16
    inp = make_inputs(B, VIEWS, hw, seed=rank)
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in inp.items()}
    res = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    out_host = torch.empty(V, H, W, NECK_CFG["out_channels"], dtype=torch.float32).pin_memory()
    forward = make_forward(model, neck, V, H, W)

    # --- end to end through the plugins: every step copies ITS inputs from pinned host memory and reads ITS result (the
    #     multi-scale feature list = level 0 of the neck, see make_forward) back to pinned host memory.  The copies run on
    #     two copy streams, double-buffered, so the H2D of step i+1 and the D2H of step i-1 overlap forward(i) (what a
    #     streaming deployment does); everything is inside the timed region, which ends only when the last D2H has landed.
    h2d_s, d2h_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    dev_in = [{k: (torch.empty_like(v, device=dev) if torch.is_tensor(v) else v) for k, v in host.items()} for _ in range(2)]
    out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]

    # the same frames as uint8 HWC camera crops (what the reference's CPU pipeline normalises, transform_3d.py:72-104)
    g8 = torch.Generator(); g8.manual_seed(100 + rank)
    host_u8 = dict(host, x=torch.randint(0, 256, (V, hw[0], hw[1], 3), generator=g8, dtype=torch.uint8).pin_memory())
    dev_in_u8 = [{k: (torch.empty_like(v, device=dev) if torch.is_tensor(v) else v) for k, v in host_u8.items()} for _ in range(2)]

    def e2e_run(steps, flush_l2, host=host, dev_in=dev_in):
        main = torch.cuda.current_stream()
        ev_in = [torch.cuda.Event() for _ in range(steps)]
        ev_fwd = [torch.cuda.Event() for _ in range(steps)]

        def stage_in(i):
            with torch.cuda.stream(h2d_s):
                if i >= 2:
                    h2d_s.wait_event(ev_fwd[i - 2])          # forward(i-2) has consumed this device buffer
                for k, v in host.items():
                    if torch.is_tensor(v):
                        dev_in[i % 2][k].copy_(v, non_blocking=True)
                ev_in[i].record(h2d_s)

        h2d_s.wait_stream(main)
        d2h_s.wait_stream(main)
        stage_in(0)
        for i in range(steps):
            if i + 1 < steps:
                stage_in(i + 1)
            if flush_l2:
                flush.zero_()
            main.wait_event(ev_in[i])
            lv0 = forward(dev_in[i % 2])
            ev_fwd[i].record(main)
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(ev_fwd[i])
                out_hosts[i % 2].copy_(lv0, non_blocking=True)
                lv0.record_stream(d2h_s)
        main.wait_stream(d2h_s)
        main.wait_stream(h2d_s)
        if gather_stream is not None:
            main.wait_stream(gather_stream)

    def e2e_timed(steps, host, dev_in):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        e2e_run(steps, True, host, dev_in)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for _ in range(max(args.warmup, 3)):
        forward(res)
    e2e_run(3, False)
    e2e_run(3, False, host_u8, dev_in_u8)
    torch.cuda.synchronize()

    with ClockSampler(local, enabled=(rank == 0)) as clk:
        l0 = L.launch_count
        total_ms, ms = timed(lambda: forward(res), args.steps)
        launches = (L.launch_count - l0) // args.steps
        e2e_ms = e2e_timed(args.steps, host, dev_in)
        e2e_u8_ms = e2e_timed(args.steps, host_u8, dev_in_u8)
    clocks = clk.summary()
    step_ms = total_ms / args.steps
    work = algorithmic_work(cfg, kind, hw, V)
    peaks = load_peaks()

    # --- roofline, method 1 (the figure reported as `achieved`): marginal time of a kernel family under the product's
    #     own conditions.  The step is re-captured as a CUDA graph with the family's C-ABI calls skipped and timed exactly
    #     like the headline; family time = step - step without the family.  Launch overlap (programmatic dependent
    #     launch), side streams and L2 state are those of the real step; events sit outside the replay, as they must.
    fam = {"gemm": ["gemm"], "attention": ["window_attention"],
           "hbm_kernels": ["layernorm_rows", "ln_gather_merge", "fast_token_update"],
           "token_kernels": ["layernorm_rows", "ln_gather_merge", "fast_token_update", "fill_pad_kv", "window_topk", "compact_rows",
                             "score_tokens", "topk_split", "motion_queries_fold", "im2col_patch16", "cast_bf16"]}
    marginal = {}
    if not args.no_roofline:
        for fname, names_ in fam.items():
            saved_ = {n: getattr(L, n) for n in names_}
            for n in names_:
                setattr(L, n, lambda *a, **k: k.get("out"))
            model._graphs = {}
            try:
                for _ in range(3):
                    forward(res)
                t_wo, _ = timed(lambda: forward(res), 10)
            finally:
                for n in names_:
                    setattr(L, n, saved_[n])
                model._graphs = {}
            marginal[fname] = max(step_ms - t_wo / 10, 1e-6)
        for _ in range(3):
            forward(res)

    # --- roofline, method 2 (per-kernel split, pessimistic): one instrumented EAGER step, CUDA events around every C-ABI
    #     launch on the launch stream.  An event between two kernels forbids their programmatic overlap, so every launch
    #     here also pays the ~3 us of prologue / launch latency that the graph replay hides: the sum exceeds the step.
    recs = []
    names = ["gemm", "window_attention", "layernorm_rows", "ln_gather_merge", "fill_pad_kv", "fill_pad_kv_rope", "compact_rows", "subln", "window_topk", "topk_split", "merge_fast_tokens",
             "fast_token_update", "motion_queries_fold", "score_tokens", "score_finish", "im2col_patch16", "mask_rows",
             "global_half_mean", "cast_bf16"]
    saved = {n: getattr(L, n) for n in names}
    kind_names = {L.EPI_LINEAR: "linear", L.EPI_QKV_ROPE: "qkv_rope", L.EPI_RESID: "resid", L.EPI_SWIGLU: "swiglu"}

    def wrap(name, fn):
        def traced(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **kw)
            e.record()
            label, fl, by = name, 0.0, 0.0
            if name == "gemm":
                A_, Bw = a[0], a[1]
                m = kw.get("M") if kw.get("M") is not None else (a[3] if len(a) > 3 and a[3] is not None else A_.shape[0])
                n_, k_ = Bw.shape
                fl = 2.0 * m * n_ * k_
                label = "gemm_" + kind_names[a[2]] + ("_lnfold" if kw.get("ln_u") is not None else "")
                # algorithmic bytes: A + B once, output once (+ fp32 residual for RESID, half-width bf16 for SWIGLU)
                by = 2.0 * m * k_ + 2.0 * n_ * k_ + {L.EPI_RESID: 8.0 * m * n_, L.EPI_SWIGLU: 1.0 * m * n_}.get(
                    a[2], (4.0 if kw.get("out_f32") else 2.0) * m * n_)
            elif name == "window_attention":              # (qkv, out, nW, seq, heads): QK^T + PV, head dim 64
                fl = 4.0 * a[2] * a[4] * a[3] * a[3] * 64
                label = "window_attention_%dx%d" % (a[2], a[3])
            elif name == "layernorm_rows":                # (x, gamma, beta, out, M, C, ...): fp32 in, bf16 out
                by = a[4] * a[5] * 6.0
            elif name == "ln_gather_merge":               # (..., nW, k, n_fast, C, eps): packed rows LN + fast rows read
                nW_, k_, nf_, C_ = a[9], a[10], a[11], a[12]
                by = nW_ * (k_ + 1) * C_ * 6.0 + nW_ * nf_ * C_ * 4.0
            elif name == "merge_fast_tokens":             # (x, fast_map, fast_score, nW, n_fast, k, C, ...)
                by = a[3] * a[4] * a[6] * 4.0
            elif name == "fast_token_update":             # (x, fast_map, packed, rep, nW, n_fast, k, C): read + write
                by = a[4] * a[5] * a[7] * 8.0
            elif name == "score_tokens":                  # (x, mask_in, A, c, V, N, C, ...)
                by = a[4] * a[5] * a[6] * 4.0
            elif name == "im2col_patch16":                # (img, out, V, Hi, Wi)
                by = a[2] * 3 * a[3] * a[4] * 6.0
            recs.append((label, s, e, fl, by))
            return r
        return traced
    for n in names:
        setattr(L, n, wrap(n, saved[n]))
    model.use_cuda_graph = False          # the timed steps replay a CUDA graph; this one launches eagerly
    torch.cuda.synchronize()
    torch.cuda._sleep(int(40e6))
    forward(res)
    torch.cuda.synchronize()
    model.use_cuda_graph = True
    for n in names:
        setattr(L, n, saved[n])
    breakdown = {}
    for label, s, e, fl, by in recs:
        b = breakdown.setdefault(label, {"n": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        b["n"] += 1; b["ms"] += s.elapsed_time(e); b["flops"] += fl; b["bytes"] += by
    for b in breakdown.values():
        b["ms"] = round(b["ms"], 4)
        fl, by = b.pop("flops"), b.pop("bytes")
        if fl:
            b["tflops"] = round(fl / (b["ms"] * 1e-3) / 1e12, 1)
        if by and not fl:
            b["gbs"] = round(by / (b["ms"] * 1e-3) / 1e9, 0)     # algorithmic bytes / event time (incl. launch gaps)
    g = [(s, e, fl) for label, s, e, fl, _ in recs if label.startswith("gemm")]
    g_ms = sum(s.elapsed_time(e) for s, e, _ in g)
    g_fl = sum(f for _, _, f in g)
    g_by = sum(by for label, _, _, _, by in recs if label.startswith("gemm"))
    a_fl = sum(fl for label, _, _, fl, _ in recs if label.startswith("window_attention"))
    ach_events = g_fl / (g_ms * 1e-3) / 1e12
    g_marg = marginal.get("gemm")
    ach = g_fl / (g_marg * 1e-3) / 1e12 if g_marg else ach_events
    traffic, traffic_src = load_ncu_traffic(args.config)
    hbm_names = ("ln_gather_merge", "fast_token_update", "merge_fast_tokens", "layernorm_rows")
    hbm_k = [v for k_, v in breakdown.items() if k_ in hbm_names and "gbs" in v]
    hbm_ms = sum(v["ms"] for v in hbm_k)
    hbm_gbs = sum(v["gbs"] * v["ms"] for v in hbm_k) / hbm_ms if hbm_ms else None
    hbm_bytes = sum(by for label, _, _, _, by in recs if label in hbm_names)
    roofline = {"bound": "tensor", "kernel": "toc3d::gemm::gemm_kernel (tcgen05 cta_group::2, all epilogues)",
                "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sust"],
                "frac_of_burst_peak": ach / peaks["tf_burst"],
                "method": ("marginal: FLOPs launched by the GEMM calls of one step / (step time - time of the same CUDA-graph step "
                           "captured with those calls skipped), CUDA events around the replays" if g_marg else
                           "per-launch CUDA events of one eager step"),
                "gemm_ms_per_step": g_marg, "gemm_share_of_step": (g_marg / step_ms) if g_marg else g_ms / step_ms,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": g_by / max(1, len(g)),
                "peak_source": peaks["src"] + " bf16_tflops_sustained", "launches_per_step": len(g),
                "avg_launch_us": (g_marg if g_marg else g_ms) * 1e3 / max(1, len(g)),
                "attention": {"ms_per_step": marginal.get("attention"), "flops": a_fl,
                              "achieved_tflops": (a_fl / (marginal["attention"] * 1e-3) / 1e12) if marginal.get("attention") else None,
                              "note": "not tensor bound: 2 MUFU (ex2) ops per 256 tensor FLOPs make the MUFU pipe the throughput floor, and thread = row on "
                                      "128-row tiles with 129- / 52- / 33-row windows leaves the launches latency-bound (DESIGN 3.2)"},
                "hbm_kernels": {"what": "LayerNorm / gather + merge (+ deferred fast-token update) / fast-token update launches: algorithmic "
                                        "bytes of one step / their marginal time (same method as `achieved`); the per-launch event "
                                        "figure of the eager step is kept as achieved_events",
                                "achieved": (hbm_bytes / (marginal["hbm_kernels"] * 1e-3) / 1e9) if marginal.get("hbm_kernels") else hbm_gbs,
                                "peak": peaks["hbm"], "unit": "GB/s",
                                "frac": ((hbm_bytes / (marginal["hbm_kernels"] * 1e-3) / 1e9) / peaks["hbm"]) if marginal.get("hbm_kernels")
                                else ((hbm_gbs / peaks["hbm"]) if hbm_gbs else None),
                                "ms_per_step": marginal.get("hbm_kernels"), "achieved_events": hbm_gbs,
                                "algorithmic_bytes_per_step": hbm_bytes},
                "token_kernels_ms_per_step": marginal.get("token_kernels"),
                "eager_event_breakdown": {"achieved_gemm_tflops": ach_events, "gemm_ms": g_ms,
                                          "note": "events between kernels forbid programmatic overlap: ~3 us per launch more than in the graph",
                                          "kernels": breakdown},
                "whole_step_tflops": work["flops"] / (step_ms * 1e-3) / 1e12}

    value = world * B * args.steps / (total_ms * 1e-3)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v))
    d2h = out_host.numel() * out_host.element_size()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": workload_config(args.config, B, world),
        "native_details": {"l2": "256 MiB buffer zeroed between timed steps (L2 flush)",
                           "residual_stream": "fp32", "gemm_operands": "bf16, fp32 accumulate",
                           "launch": "backbone + neck replayed as one CUDA graph per call",
                           "collective": ("ncclAllGather of level 0 of the feature list (NHWC fp32, %d B per rank; level 1 is its stride-2 "
                                          "subsample), %s" % (d2h, "after each forward on the launch stream" if args.serial_gather else
                                                               "on its own stream: the gather of step i overlaps forward(i+1), the last one "
                                                               "is joined and timed as a tail")) if world > 1 else None,
                           "e2e_pipeline": "per step: pinned-host inputs -> H2D, backbone + neck, feature list (level 0) -> D2H to pinned "
                                           "host; copies double-buffered on copy streams (overlap the neighbouring steps' forward); the L2 "
                                           "flush between steps is inside the e2e timed region"},
        # link_gbs = bytes moved per step / step time: when it sits at the host link's rate the e2e number is bound by the copies
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "link_gbs": (h2d + d2h) / (e2e_ms / args.steps) / 1e6},
        # row f3 (context): same pipeline fed the uint8 HWC camera crops; normalise + pad run fused in the stem kernel
        # (toc3d_preprocess_patch16_u8), so the H2D copy carries 4x fewer image bytes
        "e2e_u8_input": {"value": world * B * args.steps / (e2e_u8_ms / 1e3), "unit": UNIT,
                         "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host_u8.values() if torch.is_tensor(v)),
                         "d2h_bytes_per_step": d2h, "ms_per_step": e2e_u8_ms / args.steps},
        "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
        "clocks": clocks, "roofline": roofline,
    }
    del forward, model, neck
    torch.cuda.empty_cache()
    if not args.no_other_configs:
        # BASELINE.json's other configs through the same step definition (device-resident inputs), so that they reach the
        # driver's record: all of them at N = 1, the sharded 1600x800 config (configs[3]) at every N
        others = ["toc3d_faster", "eva_vit_l", "toc3d_fast_1600", "toc3d_faster_1600", "eva_vit_l_1600"] if world == 1 else ["toc3d_faster_1600"]
        line["other_configs"] = {n: quick_config(n, 5 if n.endswith("1600") else 10) for n in others if n != args.config}
        if world == 1 and "eva_vit_l_1600" in line["other_configs"] and "toc3d_faster_1600" in line["other_configs"]:
            line["other_configs"]["speedup_faster_1600_over_dense_1600"] = (line["other_configs"]["toc3d_faster_1600"]["value"]
                                                                            / line["other_configs"]["eva_vit_l_1600"]["value"])
    if world == 1 and args.batch == 1 and not args.no_batch4:
        # context, not the headline: the same forward with 4 samples (24 views) per launch
        m4, n4 = build_native(args.config, dev)
        res4 = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in make_inputs(4, VIEWS, hw, seed=1).items()}
        f4 = lambda: n4(list(m4(**res4).img_feats.values())) if kind == "ToC3DEVAViT" else n4(list(m4(**res4).values()))
        for _ in range(3):
            f4()
        t4, _ = timed(f4, 10)
        line["throughput_batch4"] = {"value": 4 * 10 / (t4 * 1e-3), "unit": UNIT, "ms_per_step": t4 / 10,
                                     "note": "4 samples per launch, device-resident inputs; not the headline"}
        del res4, m4, n4, f4
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference on the same box: its own modules on the B200 in PyTorch eager, fp32 and bf16
        # autocast - the number the native result should be read against; then on the host cores (cpu_baseline)
        try:
            gpu_ref = {}
            for label, ac in (("fp32", False), ("bf16_autocast", True)):
                step, rkind = reference_step_fn(args.config, VIEWS, device=dev, autocast=ac)
                for _ in range(3):
                    step()
                t_ms, _ = timed(step, 5)
                gpu_ref[label] = {"value": 5 / (t_ms * 1e-3), "unit": UNIT, "ms_per_step": t_ms / 5}
                del step
                torch.cuda.empty_cache()
            gpu_ref["what"] = ("the reference's own backbone + neck modules (.cuda(), torch %s eager, ATen/cuBLAS/cuDNN kernels), same "
                               "weights and inputs, 6 views, batch 1, L2 flushed between steps" % torch.__version__)
            line["reference_gpu_eager"] = gpu_ref
        except Exception as ex:  # reference sources not staged: say so, do not guess
            line["reference_gpu_eager"] = {"unavailable": str(ex)[:200]}
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        views = reference_views(args.config)
        step, rkind = reference_step_fn(args.config, views)
        step()
        t0 = time.perf_counter(); n = 0
        while n < 3 or (time.perf_counter() - t0 < 10 and n < 10):
            step(); n += 1
        dt = (time.perf_counter() - t0) / n
        scale = VIEWS // views
        line["cpu_baseline"] = {"value": 1.0 / (dt * scale), "unit": UNIT, "cores": cores, "kind": rkind,
                                "sample": "%d of 6 views x %d forwards (backbone + neck) after 1 warm-up, fp32, %d torch threads%s" % (
                                    views, n, cores, "" if scale == 1 else ", scaled x%d (extrapolated)" % scale)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_strong(args, world, rank, dev):
    """Strong scaling: ONE batch of `--batch` samples (6 * batch views) cut into contiguous image chunks over the ranks
    (toc3d_b200.shard.partition); every rank runs the backbone on its chunk and `last_feat`, the token masks and the
    keep / drop lists are all-gathered in global image order (ShardedBackbone over NCCL)."""
    import torch.distributed as dist
    from toc3d_b200 import shard as S
    kind, cfg, hw = CONFIGS[args.config]
    model, neck = build_native(args.config, dev)
    model.fuse_neck(None)
    Bt = args.batch
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in make_inputs(Bt, VIEWS, hw, seed=0).items()}
    sb = S.ShardedBackbone(model, views=VIEWS)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        sb(**inp)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(dev.index, enabled=(rank == 0)) as clk:
        barrier()
        for s, e in ev:
            flush.zero_()
            s.record()
            sb(**inp)
            e.record()
        barrier()
    tot = torch.tensor([sum(s.elapsed_time(e) for s, e in ev)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms = tot.item() / args.steps
    if rank == 0:
        emit({"metric": METRIC, "value": Bt / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
              "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "bf16", "data": "synthetic", "config": dict(workload_config(args.config, Bt, world), batch_total=Bt,
              parallelism="%d views of one %d-sample batch per rank (shard.partition), all-gather of last_feat + masks + index lists" % (
                  Bt * VIEWS // world, Bt)),
              "clocks": clk.summary()})
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: anything a library prints to fd 1 during the run (NCCL prints its
    version banner there) is diverted to stderr; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="toc3d_fast", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=1, help="6-view samples per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch4", action="store_true", help="skip the extra batch-4 throughput line")
    ap.add_argument("--view-groups", type=int, default=None, help="override the plugin's view_groups (streams of views)")
    ap.add_argument("--serial-gather", action="store_true", help="N > 1: all-gather on the launch stream after each forward (default: own stream, overlapped)")
    ap.add_argument("--no-roofline", action="store_true", help="skip the marginal-time roofline re-captures")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the extra lines for BASELINE.json's other configs")
    ap.add_argument("--strong", action="store_true", help="strong scaling: one --batch-sample batch sharded over the ranks (ShardedBackbone)")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
