#!/usr/bin/env python
"""bench.py — GSSD multibox hot path (match + OHNM loss fwd/bwd + Detect/NMS), images/sec.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A step = one pass of the hot path over one batch: MultiBoxLoss forward + backward (2 kernels + the
backward rescale) and Detect (1 kernel) on `--batch` images per GPU (BASELINE.json configs[1]:
batch 32, SSD300 priors P=8732, 1-5 GT boxes per image, C=2).  Weak scaling: the per-GPU batch is
fixed; every rank owns its own images, the only exchange is the 16 bytes of loss statistics per rank, stored
into the peers' memory over NVLink by the last CTA of the match kernel (no collective call).

value : inputs resident in HBM, a ring of input sets larger than L2, CUDA-graph replay of the public
        API calls, CUDA-event timing, max over ranks.
e2e   : the same step through the public Python API from pinned HOST buffers: H2D of loc / conf /
        scores / targets, kernels, D2H of the two losses and of the Detect output, every step.
roofline: the dominant kernel, timed alone with CUDA events inside this run, against
        MEASURED_PEAKS.json (hbm_gbs).   cpu_baseline: the CPU oracle (port of the reference
        algorithm, OpenMP over images) on the box's host cores, bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "GSSD images/sec (match+OHNM loss+Detect/NMS)"
UNIT = "images/s"
TOP_K, CONF_THRESH, NMS_THRESH, NEGPOS, MATCH_THRESH = 200, 0.2, 0.45, 3, 0.5
CLASS_BIAS = (0.0, -4.0)        # Detect scores = softmax(conf + bias): the "sparse-realistic" shift of BASELINE.md §3 (~3 % of priors > 0.2)
L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU (configs[1]: 32)")
    ap.add_argument("--priors", default="v2", help="prior-box config (v2: P=8732, v2_512: P=24564)")
    ap.add_argument("--gmax", type=int, default=5, help="GT boxes per image ~ U{1..gmax}")
    ap.add_argument("--no-graph", action="store_true", help="time eager API calls instead of CUDA-graph replay")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gconv", action="store_true", help="skip the tensor-core source-block measurement")
    ap.add_argument("--sweep", action="store_true", help="(default at N=1) also report the kernels' roofline at larger batches")
    ap.add_argument("--no-sweep", action="store_true", help="skip the large-batch roofline sweep")
    return ap.parse_args()


def workload_name(a):
    return "configs[1]: GSSD multibox head, batch %d/GPU, %s priors, 1-%d GT, C=2, loss fwd+bwd + Detect(thr %.1f, top_k %d, nms %.2f)" % (
        a.batch, a.priors, a.gmax, CONF_THRESH, TOP_K, NMS_THRESH)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# synthetic inputs (numpy, seeded; the same generator the tests and the golden fixtures use)
def make_inputs(a, rank, n_sets):
    from grouped_ssd_pytorch_b200 import synthetic as syn
    r = syn.rng(syn.SEED + 1000 * rank)
    sets = []
    for _ in range(n_sets):
        tg = syn.targets(r, a.batch, 1, a.gmax)
        loc = syn.loc(r, a.batch, a.P)
        conf = syn.conf_logits(r, a.batch, a.P, 2)
        x = conf.copy()
        x[..., 1] -= 4.0                                  # "sparse-realistic" Detect scores
        sets.append(dict(targets=tg, loc=loc, conf=conf, scores=syn.softmax(x)))
    return sets


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
def cpu_step_fn(a, priors_np):
    from oracle import oracle as O
    O.lib()

    def step(s):
        O.multibox_loss(s["loc"], s["conf"], priors_np, s["targets"], MATCH_THRESH, NEGPOS, (0.1, 0.2), grads=True, extras=False)
        O.detect(s["loc"], s["scores"], priors_np, 2, TOP_K, CONF_THRESH, NMS_THRESH, (0.1, 0.2))
    return step


def time_cpu(a, priors_np, sets, steps, warmup, budget_s):
    step = cpu_step_fn(a, priors_np)
    for i in range(max(1, warmup)):
        step(sets[i % len(sets)])
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        step(sets[done % len(sets)])
        done += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done, dt


def run_reference(a):
    """--impl reference: the reference algorithm (CPU oracle port, all host threads) on our config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from grouped_ssd_pytorch_b200 import config
    from oracle import oracle as O
    priors_np = O.priorbox(config.ALL[a.priors])
    a.P = priors_np.shape[0]
    sets = make_inputs(a, 0, 4)
    cores = os.cpu_count() or 1
    done, dt = time_cpu(a, priors_np, sets, a.steps, min(a.warmup, 3), budget_s=120.0)
    v = done * a.batch / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": done,
        "warmup": a.warmup, "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "batch_per_gpu": a.batch, "num_priors": a.P,
                   "note": "reference algorithm as the CPU oracle port (oracle/gssd_oracle.c, OpenMP over images); "
                           "the reference itself is Python/torch and is not present on the GPU box"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d steps of the batch-%d workload (loss fwd+bwd + Detect)" % (done, a.batch)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            load = [v for v in sm if v >= 0.5 * max(mx)] or sm        # samples taken while kernels were running
            out.update(sm_mhz=statistics.median(load), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm), samples_under_load=len(load))
        return out


# ------------------------------------------------------------------------------------------------------
def dbg(msg):
    if os.environ.get("BENCH_DEBUG"):
        sys.stderr.write("[bench r%s %.2f] %s\n" % (os.environ.get("RANK", "0"), time.perf_counter(), msg))
        sys.stderr.flush()


def run_ours(a):
    import torch
    import torch.distributed as dist
    from grouped_ssd_pytorch_b200 import _lib, config
    from grouped_ssd_pytorch_b200.layers import Detect, MultiBoxLoss, PriorBox
    from grouped_ssd_pytorch_b200.layers.box_utils import pack_target_list as pack_targets

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.require_cuda()
    dbg('init done')

    priors = PriorBox(config.ALL[a.priors]).forward(device="cuda")
    a.P = P = priors.shape[0]
    B = a.batch
    per_set = B * P * (16 + 8 + 8 + 16 + 8)              # loc, conf, scores, grad_loc, grad_conf
    n_sets = int(min(64, max(2, L2_BYTES * 1.5 // per_set + 1)))
    host = make_inputs(a, rank, n_sets)

    crit = MultiBoxLoss(2, MATCH_THRESH, True, 0, True, NEGPOS, 0.5, False, True)   # train_lesion_multiphase_v2.py:639
    dsets = []
    for s in host:
        dsets.append(dict(loc=torch.from_numpy(s["loc"]).to(dev).requires_grad_(),
                          conf=torch.from_numpy(s["conf"]).to(dev).requires_grad_(),
                          scores=torch.from_numpy(s["scores"]).to(dev),
                          # ground truth resident in HBM in its packed form (gt[sum_G,5] + row offsets): the device image
                          # of the reference's `targets` list; MultiBoxLoss takes it as is (the e2e arm packs host lists)
                          targets=pack_targets([torch.from_numpy(t) for t in s["targets"]], dev)))

    side = torch.cuda.Stream()        # Detect does not depend on the loss: it runs beside it on a second stream

    def step_device(d):
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            out = Detect.apply_logits(2, 0, TOP_K, CONF_THRESH, NMS_THRESH, d["loc"].detach(), d["conf"].detach(), priors,
                                      class_bias=CLASS_BIAS)
        d["loc"].grad = None; d["conf"].grad = None
        ll, lc = crit((d["loc"], d["conf"], priors), d["targets"])
        (ll + lc).backward()
        main.wait_stream(side)
        return ll, lc, out

    # ---- kernels per step + optional CUDA graphs -------------------------------------------------------
    for d in dsets[:2]:
        step_device(d)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    step_device(dsets[0])
    torch.cuda.synchronize()
    kernels_per_step = _lib.launch_count() - n0
    dbg('eager steps ok, kernels/step=%d' % kernels_per_step)

    graphs, use_graph = None, not a.no_graph
    if use_graph:
        try:
            cap = torch.cuda.Stream()
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                for d in dsets:
                    step_device(d)
            torch.cuda.current_stream().wait_stream(cap)
            torch.cuda.synchronize()
            graphs = []
            pool = None
            for d in dsets:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    d["result"] = step_device(d)
                pool = pool or g.pool()
                graphs.append(g)
            torch.cuda.synchronize()
            dbg('graphs captured')
        except Exception as e:                              # pragma: no cover
            sys.stderr.write("bench: CUDA-graph capture failed (%s); timing eager calls\n" % e)
            graphs, use_graph = None, False
            torch.cuda.synchronize()

    def run_step(i):
        if graphs is not None:
            graphs[i % n_sets].replay()
        else:
            step_device(dsets[i % n_sets])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: device-resident --------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    # W warm-up steps, then more in chunks until >= 0.5 s have passed so that the clocks have ramped; the
    # ranks agree on every extra chunk (the steps contain a collective, so the counts must match)
    t_w = time.perf_counter()
    for i in range(max(3, a.warmup)):
        run_step(i)
    for _ in range(200):
        torch.cuda.synchronize()
        more = torch.tensor([1.0 if time.perf_counter() - t_w < 0.5 else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(more, op=dist.ReduceOp.MAX)
        if float(more) == 0.0:
            break
        for i in range(64):
            run_step(i)
    barrier()
    dbg('warm-up done')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_launch0 = _lib.launch_count()
    e0.record()
    for i in range(a.steps):
        run_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    dbg('timed region done')
    eager_launches = _lib.launch_count() - n_launch0
    gpu_launches = kernels_per_step * a.steps if graphs is not None else eager_launches

    # ---- end to end: host buffers in, losses + detections out, every step -----------------------------------
    # The reference-facing host-buffer call (include/gssd.h gssd_pipe_*, grouped_ssd_pytorch_b200/pipeline.py): one
    # native call per step copies that step's loc / conf / scores / targets from page-locked host memory to the device,
    # runs match + loss (with gradients) + Detect, and copies the two losses and the Detect tensor back; 3 steps are in
    # flight so that step i's H2D runs beside step i-1's kernels and D2H.  The host waits for a step's results before it
    # reuses that step's buffers, and every step's copies are inside the timed region.
    from grouped_ssd_pytorch_b200.pipeline import HostPipeline
    DEPTH = 3
    pipe = HostPipeline(B, priors, num_classes=2, top_k=TOP_K, depth=DEPTH, match_thresh=MATCH_THRESH, negpos_ratio=NEGPOS,
                        conf_thresh=CONF_THRESH, nms_thresh=NMS_THRESH, max_gt_rows=B * max(a.gmax, 1),
                        detect_logits=True, class_bias=CLASS_BIAS)
    n_host = min(2 * DEPTH, max(DEPTH, n_sets))
    hbufs = []
    for i in range(n_host):
        hb = pipe.host_buffers()
        src = host[i % n_sets]
        hb.loc.copy_(torch.from_numpy(src["loc"])); hb.conf.copy_(torch.from_numpy(src["conf"]))
        hb.targets = [torch.from_numpy(t) for t in src["targets"]]
        hb.ticket = None
        hbufs.append(hb)

    def step_e2e(i):
        hb = hbufs[i % n_host]
        if hb.ticket is not None:
            pipe.wait(hb.ticket)                                 # this buffer set's previous results are on the host
        hb.ticket = pipe.submit(hb, hb.targets)

    def drain():
        for hb in hbufs:
            if hb.ticket is not None:
                pipe.wait(hb.ticket)
                hb.ticket = None

    h2d = B * P * (16 + 8) + sum(t.numel() * 4 for t in hbufs[0].targets) + 4 * (B + 1)
    d2h = 8 + B * 2 * TOP_K * 5 * 4
    e2e_steps = max(10, min(a.steps, 500))
    for i in range(2 * n_host):
        step_e2e(i)
    drain()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step_e2e(i)
    drain()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    # latency of one step with nothing else in flight (informational)
    t0 = time.perf_counter()
    for i in range(20):
        step_e2e(0)
        drain()
    e2e_serial_ms = (time.perf_counter() - t0) / 20 * 1e3
    # sanity: the pipeline returns what the device-resident public-API step returns for the same inputs
    step_e2e(0); drain()
    ll0, lc0, out0 = step_device(dsets[0])
    torch.cuda.synchronize()
    assert abs(float(hbufs[0].losses[0]) - float(ll0.detach())) <= 1e-6 * abs(float(ll0.detach())) + 1e-7, "e2e loss differs from the device-resident step"
    assert torch.equal(hbufs[0].detections, out0.cpu()), "e2e Detect output differs from the device-resident step"
    clocks = sampler.stop() if sampler else None
    dbg('e2e done')

    # ---- per-kernel durations (CUDA events on the launch stream, GPU kept busy so launches never starve) ----
    kern = time_kernels(a, lib, _lib, torch, dev, priors, dsets, pack_targets, B, P) if rank == 0 else None

    # ---- max over ranks ---------------------------------------------------------------------------------------
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        hbm, hbm_src = peaks()
        dom = max(kern, key=lambda k: k["us"])
        traffic = None
        try:                                                     # dram bytes per launch from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if int(tj.get("batch", -1)) == B and a.priors == "v2":
                traffic = tj.get(dom["name"].split(" ")[0])
        except Exception:
            traffic = None
        roof = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["gbs"], "peak": hbm, "unit": "GB/s",
                "frac": dom["gbs"] / hbm, "traffic": traffic, "peak_source": hbm_src,
                "algorithmic_bytes_per_launch": dom["bytes"], "avg_launch_us": dom["us"],
                "kernels": [{k: v for k, v in kk.items()} for kk in kern]}
        line = {
            "metric": METRIC, "value": world * B * a.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "batch_per_gpu": B, "global_batch": world * B, "num_priors": P,
                       "num_classes": 2, "parallelism": "dp%d (batch-sharded; the 16-byte loss statistics cross ranks as NVLink peer stores from the match kernel)" % world if world > 1 else "dp1",
                       "l2_policy": "ring of %d distinct input sets (%.0f MB) > L2 (126 MB)" % (n_sets, n_sets * per_set / 1e6),
                       "launch": ("cuda-graph replay of the public API calls" if graphs is not None else "eager public API calls") + "; Detect on a second stream beside the loss",
                       "kernels_per_step": kernels_per_step,
                       "targets": "value: packed ground truth resident in HBM; e2e: host list of [n_i,5] tensors packed and copied every step",
                       "detect_scores": "softmax(conf + (0,-4)) evaluated inside the Detect kernel (the CPU arm reads the same scores precomputed)"},
            "clocks": clocks,
            "e2e": {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "pipeline": "native host-buffer call (gssd_pipe_submit), 3 steps in flight: H2D beside the previous step's kernels and D2H",
                    "serial_ms_per_step": e2e_serial_ms},
            "gpu_launches": int(gpu_launches),
            "roofline": roof,
        }
        if not a.no_cpu_baseline and world == 1:                 # the CPU arm is timed on rank 0 at N = 1 only
            from oracle import oracle as O
            priors_np = priors.cpu().numpy()
            done, dt = time_cpu(a, priors_np, host, 10 ** 9, 1, a.cpu_seconds)
            line["cpu_baseline"] = {"value": done * B / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "%d steps of the batch-%d workload in %.1f s (oracle/gssd_oracle.c, OpenMP over images)" % (done, B, dt)}
        if not a.no_gconv and world == 1:
            try:
                line["gconv"] = time_gconv(a, torch, dev, B)
            except Exception as e:                               # pragma: no cover
                line["gconv"] = {"error": repr(e)}
        if (a.sweep or world == 1) and not a.no_sweep:
            line["roofline_sweep"] = sweep(a, lib, _lib, torch, dev, pack_targets)
        print(json.dumps(line), flush=True)
    if world > 1:
        # graphs that captured NCCL work must be gone before the communicator is torn down; a watchdog
        # guarantees the ranks exit even if the teardown stalls (results are already printed)
        dist.barrier()
        graphs = None
        for d in dsets:
            d.pop("result", None)
        import gc
        gc.collect()
        torch.cuda.synchronize()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        try:
            dist.destroy_process_group()
        finally:
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)


def time_kernels(a, lib, _lib, torch, dev, priors, dsets, pack_targets, B, P, iters=None):
    """Average duration of each of our kernels, timed alone: a spin kernel keeps the GPU busy while the
    CPU enqueues [event, kernel, event] x n, so no launch ever waits for the host."""
    n_sets = len(dsets)
    iters = iters or max(20, min(100, 4 * n_sets))
    st = _lib.stream()
    packed = [tuple(pack_targets(d["targets"], dev)) for d in dsets]
    tags = torch.empty((B, P), dtype=torch.int16, device=dev)
    stats = torch.empty((16 + 4 * B,), dtype=torch.uint8, device=dev)
    losses = torch.empty((2,), dtype=torch.float32, device=dev)
    gl = [torch.empty((B, P, 4), dtype=torch.float32, device=dev) for _ in range(n_sets)]
    gc = [torch.empty((B, P, 2), dtype=torch.float32, device=dev) for _ in range(n_sets)]
    wsb = lib.gssd_workspace_bytes(_lib.WS_LOSS, B, P, 2, 1, 0)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    out = torch.empty((B, 2, TOP_K, 5), dtype=torch.float32, device=dev)

    import ctypes
    bias = (ctypes.c_float * 2)(*CLASS_BIAS)

    def k_match(i):
        gt, off, sg, gm = packed[i]
        _lib.check(lib.gssd_mbox_match(priors.data_ptr(), P, dsets[i]["conf"].data_ptr(), 2, gt.data_ptr(), off.data_ptr(),
                                       B, sg, gm, MATCH_THRESH, tags.data_ptr(), stats.data_ptr(), st))

    def k_loss(i):
        gt, off, sg, gm = packed[i]
        _lib.check(lib.gssd_mbox_loss(dsets[i]["loc"].data_ptr(), dsets[i]["conf"].data_ptr(), priors.data_ptr(), B, P, 2,
                                      gt.data_ptr(), off.data_ptr(), sg, gm, tags.data_ptr(), stats.data_ptr(), None, 0,
                                      NEGPOS, 0.1, 0.2, losses.data_ptr(), gl[i].data_ptr(), gc[i].data_ptr(), None, None,
                                      ws.data_ptr(), wsb, st))

    def k_det(i):
        _lib.check(lib.gssd_detect_logits(dsets[i]["loc"].data_ptr(), dsets[i]["conf"].data_ptr(), bias, priors.data_ptr(), B, P, 2,
                                          TOP_K, CONF_THRESH, NMS_THRESH, 0.1, 0.2, out.data_ptr(), None, None, st))

    res = []
    specs = [("gssd_mbox_match (match_kernel: IoU sweep + conf max)", k_match, B * P * (16 + 8 + 2)),
             ("gssd_mbox_loss (loss_kernel: encode + smooth-L1 + OHNM select + CE + grads)", k_loss, B * P * 64),
             ("gssd_detect_logits (detect_kernel: softmax + threshold + top-k + decode + NMS)", k_det, B * (P * 40 + 8000))]
    k_match(0)
    for name, fn, nbytes in specs:
        for i in range(3):
            fn(i % n_sets)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        torch.cuda._sleep(int(2.5e7))                    # ~13 ms of GPU spin: the host runs ahead
        for i in range(iters):
            ev[i][0].record()
            fn(i % n_sets)
            ev[i][1].record()
        torch.cuda.synchronize()
        us = statistics.mean(e[0].elapsed_time(e[1]) for e in ev) * 1e3
        res.append({"name": name, "us": us, "bytes": nbytes, "gbs": nbytes / us / 1e3})
    return res


def time_gconv(a, torch, dev, B):
    """the source-1 block of configs[1] (batch B, 38x38, 512 channels) as three launches of the tcgen05/TMEM
    implicit-GEMM kernel — vgg.30 grouped 3x3 (groups 4) -> fuse_11 1x1 (+deferred L2Norm) -> loc/conf 3x3 heads —
    each timed alone with CUDA events over a ring of inputs larger than L2, against the measured bf16 GEMM peak."""
    import torch.nn as nn
    from grouped_ssd_pytorch_b200.layers.modules.source_block import PM, _Conv, conv_igemm
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst: kernels timed alone)"
    except Exception:
        peak, src = 1590.0, "fallback (B200_PROFILING.md)"
    torch.manual_seed(1111)
    C, HW, A, NC = 512, 38, 4, 2
    g = _Conv(nn.Conv2d(C, C, 3, padding=1, groups=4).to(dev), 4, dev=dev)
    f = _Conv(nn.Conv2d(C, C, 1).to(dev), 1, dev=dev)
    h = _Conv(nn.Conv2d(C, A * 4, 3, padding=1).to(dev), 1, extra=nn.Conv2d(C, A * NC, 3, padding=1).to(dev), dev=dev)
    n_sets = 3                                                       # 3 x 52 MB activations > L2
    xs = [PM.from_nchw(torch.relu(torch.randn(B, C, HW, HW, device=dev))) for _ in range(n_sets)]
    P = HW * HW * A
    loc, conf = torch.empty(B, P, 4, device=dev), torch.empty(B, P, NC, device=dev)
    ss = torch.empty((xs[0].rows,), dtype=torch.float32, device=dev)
    ys = [conv_igemm(x, g, relu=True, shift=g.bias, row_ss_out=ss) for x in xs]
    zs = [conv_igemm(y, f, relu=True, shift=f.bias, row_ss_in=ss, l2_eps=1e-10) for y in ys]
    specs = [("vgg.30 grouped 3x3 (groups 4, 512->512) + bias/BN-fold + ReLU + L2Norm row sums",
              lambda i: conv_igemm(xs[i], g, relu=True, shift=g.bias, row_ss_out=ss), 2.0 * B * HW * HW * C * (C // 4) * 9),
             ("fuse_11 1x1 (512->512) + deferred L2Norm + bias/BN-fold + ReLU",
              lambda i: conv_igemm(ys[i], f, relu=True, shift=f.bias, row_ss_in=ss, l2_eps=1e-10), 2.0 * B * HW * HW * C * C),
             ("loc.0 + conf.0 3x3 heads (512->24) + NHWC flatten/concat into loc/conf",
              lambda i: conv_igemm(zs[i], h, relu=False, shift=h.bias, head=(loc, conf, A, NC, 0, P)), 2.0 * B * HW * HW * (A * (4 + NC)) * C * 9)]
    rows, iters = [], 24
    for name, fn, flops in specs:
        for i in range(3):
            fn(i % n_sets)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        torch.cuda._sleep(int(2.5e7))
        for i in range(iters):
            ev[i][0].record()
            fn(i % n_sets)
            ev[i][1].record()
        torch.cuda.synchronize()
        us = statistics.mean(e[0].elapsed_time(e[1]) for e in ev) * 1e3
        rows.append({"name": name, "us": us, "flops": flops, "tflops": flops / us / 1e6, "frac": flops / us / 1e6 / peak})
    tot_us, tot_fl = sum(r["us"] for r in rows), sum(r["flops"] for r in rows)
    return {"bound": "tensor", "workload": "configs[1] source 1: batch %d x 38x38 x 512, grouped conv -> fuse 1x1 -> heads, bf16 in / fp32 accumulate" % B,
            "achieved": tot_fl / tot_us / 1e6, "peak": peak, "unit": "TFLOP/s", "frac": tot_fl / tot_us / 1e6 / peak, "peak_source": src,
            "images_per_s": B / (tot_us * 1e-6), "total_us": tot_us,
            "flops_note": "algorithmic flops of the 38x38 interior; the kernel also computes the 1-pixel border (40x40 rows, +10.8%)",
            "kernels": rows}


def sweep(a, lib, _lib, torch, dev, pack_targets):
    """The same three kernels at batches that fill the machine (configs[3]/[4] shapes, one GPU's share and
    beyond): where the HBM roofline fraction of each kernel saturates."""
    from grouped_ssd_pytorch_b200 import config, synthetic as syn
    from grouped_ssd_pytorch_b200.layers import PriorBox
    hbm, _ = peaks()
    rows = []
    for pname, B, gmax in (("v2", 256, 5), ("v2", 1024, 5), ("v2_512", 64, 32), ("v2_512", 512, 32)):
        pri = PriorBox(config.ALL[pname]).forward(device="cuda")
        P = pri.shape[0]
        per_set = B * P * 56
        n_sets = int(min(8, max(2, 200e6 // per_set + 1)))
        r = syn.rng(7)
        dsets = []
        for _ in range(n_sets):
            conf = torch.randn((B, P, 2), device=dev)
            dsets.append(dict(loc=torch.randn((B, P, 4), device=dev) * 0.5, conf=conf,
                              scores=torch.softmax(conf + torch.tensor([0.0, -4.0], device=dev), -1),
                              targets=[torch.from_numpy(t).to(dev) for t in syn.targets(r, B, 1, gmax)]))
        for k in time_kernels(a, lib, _lib, torch, dev, pri, dsets, pack_targets, B, P, iters=12):
            rows.append({"priors": pname, "batch": B, "gmax": gmax, "kernel": k["name"].split(" ")[0], "us": k["us"],
                         "gbs": k["gbs"], "frac": k["gbs"] / hbm})
        del dsets
        torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
