#!/usr/bin/env python
"""bench.py — GSSD multibox hot path (match + OHNM loss fwd/bwd + Detect/NMS), images/sec.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A step = one pass of the hot path over one batch.  `--config` picks the BASELINE.json workload:
  1 (default) configs[1]: MultiBoxLoss forward + backward and Detect on 32 images per GPU, SSD300 priors (P=8732),
              1-5 GT boxes per image, C=2.  Weak scaling (the per-GPU batch is fixed).
  3           configs[3]: Detect/NMS only, global batch 256 sharded over the N GPUs (strong scaling), conf_thresh 0.2,
              top_k 200, nms 0.45.
  4           configs[4]: matching + OHNM loss forward + backward only, SSD512 priors (P=24564), up to 32 GT boxes per
              image, global batch 512 sharded over the N GPUs (strong scaling).
Every rank owns its own images; the only exchange is the 16 bytes of loss statistics per rank, stored into the peers'
memory over NVLink by the kernels themselves (no collective call).  At N > 1 rank 0 also gathers the global batch once,
runs it through the single-GPU path and asserts that the sharded result is the same (`parity_check` in the JSON line).

value : inputs resident in HBM, a ring of input sets larger than L2, CUDA-graph replay of the public
        API calls, CUDA-event timing, max over ranks.
e2e   : the same step through the reference-facing host-buffer call (gssd_pipe_submit) from pinned HOST buffers: H2D of
        loc / conf / targets, kernels, D2H of the two losses and of the Detect output, every step.
Both timed regions run a whole number of rounds of `--steps` steps, as many as it takes to reach `--min-seconds` (0.5 s),
so that a short `--steps` cannot turn the number into a measurement of launch jitter; `steps` in the JSON line is what
was timed.
roofline: the dominant kernel, timed alone with CUDA events inside this run, against
        MEASURED_PEAKS.json (hbm_gbs).   cpu_baseline: the CPU oracle (port of the reference
        algorithm, OpenMP over images) on the box's host cores, bounded sample.
"""
import argparse
import json
import math
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
PRESETS = {
    1: dict(mode="both", scaling="weak", batch=32, priors="v2", gmax=5),
    3: dict(mode="detect", scaling="strong", global_batch=256, priors="v2", gmax=5),
    4: dict(mode="loss", scaling="strong", global_batch=512, priors="v2_512", gmax=32),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(PRESETS), help="BASELINE.json configs[i] preset")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (overrides the preset)")
    ap.add_argument("--priors", default=None, help="prior-box config (v2: P=8732, v2_512: P=24564; overrides the preset)")
    ap.add_argument("--gmax", type=int, default=None, help="GT boxes per image ~ U{1..gmax} (overrides the preset)")
    ap.add_argument("--min-seconds", type=float, default=0.5, help="minimum duration of each timed region")
    ap.add_argument("--no-graph", action="store_true", help="time eager API calls instead of CUDA-graph replay")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gconv", action="store_true", help="skip the tensor-core source-block measurement")
    ap.add_argument("--sweep", action="store_true", help="(default at N=1, config 1) also report the kernels' roofline at larger batches")
    ap.add_argument("--no-sweep", action="store_true", help="skip the large-batch roofline sweep")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the gathered-batch check against the single-GPU path")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    resolve(a, max(1, max(world, a.gpus) if a.impl == "reference" else world))
    return a


def resolve(a, world):
    """fill the workload fields from the preset; a.batch is PER GPU from here on"""
    pre = PRESETS[a.config]
    a.mode, a.scaling = pre["mode"], pre["scaling"]
    a.priors = a.priors or pre["priors"]
    a.gmax = a.gmax or pre["gmax"]
    if a.batch is None:
        a.batch = pre["batch"] if a.scaling == "weak" else max(1, pre["global_batch"] // world)
    a.world = world
    a.global_batch = a.batch * world


def workload_name(a):
    what = {"both": "loss fwd+bwd + Detect(thr %.1f, top_k %d, nms %.2f)" % (CONF_THRESH, TOP_K, NMS_THRESH),
            "detect": "Detect only (thr %.1f, top_k %d, nms %.2f)" % (CONF_THRESH, TOP_K, NMS_THRESH),
            "loss": "matching + OHNM loss fwd+bwd only"}[a.mode]
    return "configs[%d]: GSSD multibox head, batch %d/GPU, %s priors, 1-%d GT, C=2, %s" % (a.config, a.batch, a.priors, a.gmax, what)


def config_dict(a, P):
    """the `config` object: identical (keys and values) in our arm and in the reference arm of the same command line"""
    return {"workload": workload_name(a), "preset": a.config, "mode": a.mode, "batch_per_gpu": a.batch,
            "global_batch": a.global_batch, "num_priors": P, "num_classes": 2, "gt_per_image": "U{1..%d}" % a.gmax,
            "parallelism": "dp%d" % a.world, "scaling": a.scaling,
            "l2_policy": "ring of distinct input sets larger than L2 (126 MB)"}


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
        sets.append(dict(targets=tg, loc=loc, conf=conf))
    return sets


def detect_scores(conf):
    """softmax(conf + CLASS_BIAS): the "sparse-realistic" Detect scores the CPU arm reads precomputed"""
    from grouped_ssd_pytorch_b200 import synthetic as syn
    x = conf.copy()
    x[..., 0] += np.float32(CLASS_BIAS[0]); x[..., 1] += np.float32(CLASS_BIAS[1])
    return syn.softmax(x)


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
def cpu_threads():
    """all the host threads this process may use (cgroup / affinity aware), set explicitly: torchrun exports
    OMP_NUM_THREADS=1, which would silently make the oracle single-threaded"""
    from oracle import oracle as O
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return O.set_threads(n)


def cpu_step_fn(a, priors_np):
    from oracle import oracle as O
    O.lib()

    def step(s):
        if a.mode in ("both", "loss"):
            O.multibox_loss(s["loc"], s["conf"], priors_np, s["targets"], MATCH_THRESH, NEGPOS, (0.1, 0.2), grads=True, extras=False)
        if a.mode in ("both", "detect"):
            if "scores" not in s:
                s["scores"] = detect_scores(s["conf"])
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


def cpu_sample_text(a, done, dt):
    what = {"both": "loss fwd+bwd + Detect", "detect": "Detect", "loss": "loss fwd+bwd"}[a.mode]
    return "%d steps of the batch-%d workload (%s) in %.1f s (oracle/gssd_oracle.c, OpenMP over images)" % (done, a.batch, what, dt)


def run_reference(a):
    """--impl reference: the reference algorithm (CPU oracle port, all host threads) on our config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from grouped_ssd_pytorch_b200 import config
    from oracle import oracle as O
    priors_np = O.priorbox(config.ALL[a.priors])
    a.P = priors_np.shape[0]
    sets = make_inputs(a, 0, 4 if a.batch * a.P < 4e6 else 2)
    threads = cpu_threads()
    done, dt = time_cpu(a, priors_np, sets, a.steps, min(a.warmup, 3), budget_s=60.0)
    v = done * a.batch / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": done,
        "warmup": a.warmup, "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(a, a.P),
        "details": {"note": "reference algorithm as the CPU oracle port (oracle/gssd_oracle.c, OpenMP over images) on the host "
                            "cores of the box; the reference itself is Python/torch and is not present on the GPU box. "
                            "Each step is one batch_per_gpu-sized batch of the workload (a bounded sample)",
                    "omp_threads": threads, "os_cpu_count": os.cpu_count()},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample_text(a, done, dt)},
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


def pin_to_gpu_numa(torch, local):
    """bind this rank's host threads (and therefore its page-locked buffers, first touch) to the CPUs next to its GPU:
    at N = 8 the ranks otherwise share whichever socket the launcher started them on.  Returns a description."""
    try:
        pr = torch.cuda.get_device_properties(local)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(path) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return "cpus %s (local to the GPU)" % txt
        return "unchanged (%d cpus allowed, GPU-local list %s)" % (len(allowed), txt)
    except Exception as e:                                       # pragma: no cover
        return "unavailable (%s)" % type(e).__name__


def nrank_parity(a, torch, dist, dev, rank, world, priors, dsets, MultiBoxLoss, Detect):
    """N > 1: the reference computes the loss on the DataParallel-gathered global batch (train_lesion_multiphase_v2.py:242-246
    -> multibox_loss.py:117, box_utils.py:167).  Every rank runs its shard of input set 0 through the sharded path (statistics
    exchanged between the ranks); rank 0 gathers inputs and results, runs the whole batch through the single-GPU path and
    compares: sum over ranks of the losses within 1e-6 relative, positive / hard-negative masks and Detect output bit-equal."""
    d = dsets[0]
    res = {"checked": False}
    out_l = {}
    if a.mode in ("both", "loss"):
        crit = MultiBoxLoss(2, MATCH_THRESH, True, 0, True, NEGPOS, 0.5, False, True)
        crit.keep_masks = True
        loc = d["loc"].detach().clone().requires_grad_()
        conf = d["conf"].detach().clone().requires_grad_()
        ll, lc = crit((loc, conf, priors), d["targets"])
        (ll + lc).backward()
        out_l = dict(losses=torch.stack([ll.detach(), lc.detach()]), pos=crit.last_masks["pos"].clone(),
                     neg=crit.last_masks["neg"].clone(), gl=loc.grad, gc=conf.grad)
    det = None
    if a.mode in ("both", "detect"):
        det = Detect.apply_logits(2, 0, TOP_K, CONF_THRESH, NMS_THRESH, d["loc"].detach(), d["conf"].detach(), priors, class_bias=CLASS_BIAS)
    torch.cuda.synchronize()

    def gather(t):
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t.contiguous(), bufs, dst=0)
        return bufs

    g_loc, g_conf = gather(d["loc"].detach()), gather(d["conf"].detach())
    gt, gt_off = d["targets"].gt, d["targets"].gt_off
    # packed ground truth: rows differ per rank -> pad to the global maximum
    n_rows = torch.tensor([gt.shape[0]], device=dev)
    all_rows = [torch.zeros_like(n_rows) for _ in range(world)]
    dist.all_gather(all_rows, n_rows)
    max_rows = int(max(int(x) for x in all_rows))
    gt_pad = torch.zeros((max_rows, 5), device=dev); gt_pad[:gt.shape[0]] = gt
    g_gt, g_off = gather(gt_pad), gather(gt_off)
    g_out = {k: gather(v) for k, v in out_l.items()}
    g_det = gather(det) if det is not None else None
    if rank == 0:
        from grouped_ssd_pytorch_b200.layers.box_utils import pack_target_list
        LOC, CONF = torch.cat(g_loc), torch.cat(g_conf)
        tg = []
        for r in range(world):
            off = g_off[r].cpu().tolist()
            tg += [g_gt[r][off[i]:off[i + 1]] for i in range(len(off) - 1)]
        res = {"checked": True, "global_batch": int(LOC.shape[0])}
        if out_l:
            crit1 = MultiBoxLoss(2, MATCH_THRESH, True, 0, True, NEGPOS, 0.5, False, True)
            crit1.process_group = False                          # single-GPU path: statistics of the whole batch, no exchange
            crit1.keep_masks = True
            L1, C1 = LOC.clone().requires_grad_(), CONF.clone().requires_grad_()
            ll1, lc1 = crit1((L1, C1, priors), pack_target_list(tg, dev))
            (ll1 + lc1).backward()
            tot = torch.stack(g_out["losses"]).double().sum(0)
            ref = torch.stack([ll1.detach(), lc1.detach()]).double()
            rel = ((tot - ref).abs() / ref.abs()).max().item()
            pos_eq = bool(torch.equal(torch.cat(g_out["pos"]), crit1.last_masks["pos"]))
            neg_eq = bool(torch.equal(torch.cat(g_out["neg"]), crit1.last_masks["neg"]))
            g_err = max((torch.cat(g_out["gl"]) - L1.grad).abs().max().item(), (torch.cat(g_out["gc"]) - C1.grad).abs().max().item())
            g_scale = max(L1.grad.abs().max().item(), C1.grad.abs().max().item())
            res.update(loss_rel_err=rel, pos_mask_equal=pos_eq, neg_mask_equal=neg_eq, grad_max_abs_err=g_err, grad_scale=g_scale,
                       num_pos=int(crit1.last_masks["num_pos"].sum().item()))
            assert rel <= 1e-6, "N-rank loss differs from the single-GPU loss of the gathered batch: rel %.3e" % rel
            assert pos_eq and neg_eq, "N-rank positive / hard-negative masks differ from the single-GPU path"
            assert g_err <= 1e-6 * g_scale + 1e-12, "N-rank gradients differ from the single-GPU path"
        if g_det is not None:
            det1 = Detect.apply_logits(2, 0, TOP_K, CONF_THRESH, NMS_THRESH, LOC, CONF, priors, class_bias=CLASS_BIAS)
            det_eq = bool(torch.equal(torch.cat(g_det), det1))
            res.update(detect_equal=det_eq)
            assert det_eq, "N-rank Detect output differs from the single-GPU path"
        torch.cuda.synchronize()
    dist.barrier()
    return res


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
    numa = pin_to_gpu_numa(torch, local) if world > 1 else "not applied (1 rank)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.require_cuda()
    dbg('init done')
    do_loss, do_det = a.mode in ("both", "loss"), a.mode in ("both", "detect")

    priors = PriorBox(config.ALL[a.priors]).forward(device="cuda")
    a.P = P = priors.shape[0]
    B = a.batch
    per_set = B * P * ((16 + 8) + (16 + 8 if do_loss else 0))   # loc, conf (+ grad_loc, grad_conf)
    n_sets = int(min(64, max(2, L2_BYTES * 1.5 // per_set + 1)))
    host = make_inputs(a, rank, n_sets)

    crit = MultiBoxLoss(2, MATCH_THRESH, True, 0, True, NEGPOS, 0.5, False, True)   # train_lesion_multiphase_v2.py:639
    dsets = []
    for s in host:
        dsets.append(dict(loc=torch.from_numpy(s["loc"]).to(dev).requires_grad_(),
                          conf=torch.from_numpy(s["conf"]).to(dev).requires_grad_(),
                          # ground truth resident in HBM in its packed form (gt[sum_G,5] + row offsets): the device image
                          # of the reference's `targets` list; MultiBoxLoss takes it as is (the e2e arm packs host lists)
                          targets=pack_targets([torch.from_numpy(t) for t in s["targets"]], dev)))

    side = torch.cuda.Stream()        # Detect does not depend on the loss: it runs beside it on a second stream

    def step_device(d):
        main = torch.cuda.current_stream()
        out = ll = lc = None
        if do_det and do_loss:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                out = Detect.apply_logits(2, 0, TOP_K, CONF_THRESH, NMS_THRESH, d["loc"].detach(), d["conf"].detach(), priors,
                                          class_bias=CLASS_BIAS)
        elif do_det:
            out = Detect.apply_logits(2, 0, TOP_K, CONF_THRESH, NMS_THRESH, d["loc"].detach(), d["conf"].detach(), priors,
                                      class_bias=CLASS_BIAS)
        if do_loss:
            d["loc"].grad = None; d["conf"].grad = None
            ll, lc = crit((d["loc"], d["conf"], priors), d["targets"])
            (ll + lc).backward()
        if do_det and do_loss:
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

    # ---- N > 1: the sharded step against the single-GPU path on the gathered batch ---------------------------
    parity = None
    if world > 1 and not a.no_parity_check:
        parity = nrank_parity(a, torch, dist, dev, rank, world, priors, dsets, MultiBoxLoss, Detect)
        dbg('parity check done: %s' % parity)

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
            # ONE graph holds one step on every input set of the ring, in order, on the stream(s) the API calls use: a replay is
            # n_sets consecutive steps.  (A graph per step costs a graph launch per 33 us of work: ~10 us of every step.)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for d in dsets:
                    d["result"] = step_device(d)
            graphs = g
            torch.cuda.synchronize()
            dbg('graph captured: %d steps' % n_sets)
        except Exception as e:                              # pragma: no cover
            sys.stderr.write("bench: CUDA-graph capture failed (%s); timing eager calls\n" % e)
            graphs, use_graph = None, False
            torch.cuda.synchronize()

    def run_steps(n):
        """at least n steps, whole ring passes when replaying the graph; returns how many ran"""
        if graphs is not None:
            reps = (n + n_sets - 1) // n_sets
            for _ in range(reps):
                graphs.replay()
            return reps * n_sets
        for i in range(n):
            step_device(dsets[i % n_sets])
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def agree_max(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- timed region: device-resident --------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    # W warm-up steps, then more in chunks until >= 0.5 s have passed so that the clocks have ramped; the
    # ranks agree on every extra chunk (the steps contain an exchange, so the counts must match)
    K = max(1, a.steps)
    t_w = time.perf_counter()
    run_steps(max(3, a.warmup))
    for _ in range(200):
        torch.cuda.synchronize()
        if agree_max(1.0 if time.perf_counter() - t_w < 0.5 else 0.0) == 0.0:
            break
        run_steps(64)
    # one untimed round of K steps sizes the timed region: whole rounds of K steps, at least --min-seconds of them
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k_done = run_steps(K)
    e1.record()
    barrier()
    round_ms = agree_max(e0.elapsed_time(e1)) * K / k_done
    rounds = int(max(1, min(10000, math.ceil(a.min_seconds * 1e3 / max(round_ms, 1e-3)))))
    dbg('warm-up done; %d rounds of %d steps' % (rounds, K))
    n_launch0 = _lib.launch_count()
    e0.record()
    steps = run_steps(rounds * K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    dbg('timed region done')
    eager_launches = _lib.launch_count() - n_launch0
    gpu_launches = kernels_per_step * steps if graphs is not None else eager_launches

    # ---- end to end: host buffers in, losses + detections out, every step -----------------------------------
    # The reference-facing host-buffer call (include/gssd.h gssd_pipe_*, grouped_ssd_pytorch_b200/pipeline.py): one
    # native call per step copies that step's loc / conf / targets from page-locked host memory to the device,
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
        hb.targets = [torch.from_numpy(t) for t in src["targets"]] if do_loss else None
        hb.ticket = None
        hbufs.append(hb)

    def step_e2e(i):
        hb = hbufs[i % n_host]
        if hb.ticket is not None:
            pipe.wait(hb.ticket)                                 # this buffer set's previous results are on the host
        hb.ticket = pipe.submit(hb, hb.targets, detect=do_det)

    def drain():
        for hb in hbufs:
            if hb.ticket is not None:
                pipe.wait(hb.ticket)
                hb.ticket = None

    h2d = B * P * (16 + 8) + ((sum(t.numel() * 4 for t in hbufs[0].targets) + 4 * (B + 1)) if do_loss else 0)
    d2h = (8 if do_loss else 0) + (B * 2 * TOP_K * 5 * 4 if do_det else 0)
    for i in range(2 * n_host):
        step_e2e(i)
    drain()
    barrier()
    K2 = max(10, min(K, 500))
    t0 = time.perf_counter()
    for i in range(K2):
        step_e2e(i)
    drain()
    torch.cuda.synchronize()
    round_s = agree_max(time.perf_counter() - t0)
    e2e_steps = int(max(1, min(2000, math.ceil(a.min_seconds / max(round_s, 1e-6))))) * K2
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
    if do_loss:
        assert abs(float(hbufs[0].losses[0]) - float(ll0.detach())) <= 1e-6 * abs(float(ll0.detach())) + 1e-7, "e2e loss differs from the device-resident step"
    if do_det:
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
            "metric": METRIC, "value": world * B * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(3, a.warmup), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(a, P),
            "details": {"steps_arg": K, "rounds_of_steps_arg": rounds,
                        "exchange": "batch-sharded; the 16-byte loss statistics cross ranks as NVLink peer stores issued by the kernels" if world > 1 else "none (1 rank)",
                        "l2_ring": "%d distinct input sets (%.0f MB)" % (n_sets, n_sets * per_set / 1e6),
                        "launch": ("cuda-graph replay of the public API calls, one graph = one step on each of the ring's input sets" if graphs is not None else "eager public API calls") + ("; Detect on a second stream beside the loss" if a.mode == "both" else ""),
                        "kernels_per_step": kernels_per_step, "host_numa": numa,
                        "targets": "value: packed ground truth resident in HBM; e2e: host list of [n_i,5] tensors packed and copied every step",
                        "detect_scores": "softmax(conf + (0,-4)) evaluated inside the Detect kernel (the CPU arm reads the same scores precomputed)"},
            "clocks": clocks,
            "e2e": {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "h2d_gbs_per_rank": h2d * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                    "pipeline": "native host-buffer call (gssd_pipe_submit), 3 steps in flight: H2D beside the previous step's kernels and D2H",
                    "serial_ms_per_step": e2e_serial_ms},
            "gpu_launches": int(gpu_launches),
            "roofline": roof,
        }
        if parity is not None:
            line["parity_check"] = parity
        if not a.no_cpu_baseline and world == 1:                 # the CPU arm is timed on rank 0 at N = 1 only
            priors_np = priors.cpu().numpy()
            threads = cpu_threads()
            done, dt = time_cpu(a, priors_np, host, 10 ** 9, 1, a.cpu_seconds)
            line["cpu_baseline"] = {"value": done * B / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": cpu_sample_text(a, done, dt)}
        if not a.no_gconv and world == 1 and a.config == 1:
            try:
                line["gconv"] = time_gconv(a, torch, dev, B)
            except Exception as e:                               # pragma: no cover
                line["gconv"] = {"error": repr(e)}
        if not a.no_gconv and world == 1 and a.config == 1:
            try:
                line["model_step"] = time_model_step(a, torch, dev, B)
            except Exception as e:                               # pragma: no cover
                line["model_step"] = {"error": repr(e)}
        if not a.no_gconv and world == 1 and a.config == 1:
            try:
                line["gssdpp"] = time_gssdpp(torch, dev)
            except Exception as e:                               # pragma: no cover
                line["gssdpp"] = {"error": repr(e)}
        if (a.sweep or (world == 1 and a.config == 1)) and not a.no_sweep:
            try:
                line["roofline_sweep"] = sweep(a, lib, _lib, torch, dev, pack_targets)
            except Exception as e:                               # pragma: no cover  (informational section: the line still prints)
                line["roofline_sweep"] = {"error": repr(e)}
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

    g_max_all = max(pk[3] for pk in packed)
    fused_ok = lib.gssd_mbox_fused_supported(B, P, 2, g_max_all) == 1
    fstate = torch.zeros((int(lib.gssd_fused_state_bytes()),), dtype=torch.uint8, device=dev)

    def k_fused(i):
        gt, off, sg, gm = packed[i]
        _lib.check(lib.gssd_mbox_loss_fused(dsets[i]["loc"].data_ptr(), dsets[i]["conf"].data_ptr(), priors.data_ptr(), B, P, 2,
                                            gt.data_ptr(), off.data_ptr(), sg, gm, MATCH_THRESH, NEGPOS, 0.1, 0.2, fstate.data_ptr(), None,
                                            losses.data_ptr(), gl[i].data_ptr(), gc[i].data_ptr(), None, None, None,
                                            ws.data_ptr(), wsb, st))

    def k_det(i):
        _lib.check(lib.gssd_detect_logits(dsets[i]["loc"].data_ptr(), dsets[i]["conf"].data_ptr(), bias, priors.data_ptr(), B, P, 2,
                                          TOP_K, CONF_THRESH, NMS_THRESH, 0.1, 0.2, out.data_ptr(), None, None, st))

    res = []
    mode = getattr(a, "mode", "both")
    specs = []
    if mode in ("both", "loss"):
        if fused_ok:        # the batch has a one-launch form: that is what MultiBoxLoss runs (fused fwd+bwd = 64 B per prior)
            specs += [("gssd_mbox_loss_fused (fused_kernel: match + encode + smooth-L1 + OHNM select + CE + grads, one launch)", k_fused, B * P * 64)]
        else:
            specs += [("gssd_mbox_match (match_kernel: IoU sweep + conf max)", k_match, B * P * (16 + 8 + 2)),
                      ("gssd_mbox_loss (loss_kernel: encode + smooth-L1 + OHNM select + CE + grads)", k_loss, B * P * 64)]
    if mode in ("both", "detect"):
        specs += [("gssd_detect_logits (detect_kernel: softmax + threshold + top-k + decode + NMS)", k_det, B * (P * 40 + 8000))]
    k_match(0)
    for name, fn, nbytes in specs:
        for i in range(3):
            fn(i % n_sets)
        torch.cuda.synchronize()
        # `iters` back-to-back launches between ONE pair of events (an event pair per launch adds ~4 us of its own to a 25 us
        # kernel), behind ~13 ms of GPU spin so that the host is ahead and no launch waits for it
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(2.5e7))
        e0.record()
        for i in range(iters):
            fn(i % n_sets)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
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


def time_gssdpp(torch, dev, B=4):
    """SURVEY f4 at configs[2]'s per-GPU share (4 images): GSSD++'s deformable convolution (1024 -> 512 channels, 38 x 38, 4
    deformable groups) and the attention core of its Self_Attn blocks (C = 512, 38 x 38) on this library's kernels against the
    operators the reference runs them on here — torchvision's fp32 deform_conv2d (stand-in of the un-vendored dcn_v2) and
    torch's bmm + softmax + bmm — forward and forward + backward, CUDA events."""
    from torchvision.ops import deform_conv2d
    from grouped_ssd_pytorch_b200.layers import dcn_v2_custom as D, self_attn as S
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False

    def timed(fn, iters=10, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    def fwd_and_step(fn, leaves, gout):
        def step():
            for t in leaves:
                t.grad = None
            fn().backward(gout)
        with torch.no_grad():
            f = timed(fn)
        return f, timed(step)

    torch.manual_seed(1111)
    C, O, H, dg = 1024, 512, 38, 4
    x = torch.randn(B, C, H, H, device=dev, requires_grad=True)
    w = (torch.randn(O, C, 3, 3, device=dev) / (9 * C) ** 0.5).requires_grad_(True)
    b = torch.zeros(O, device=dev, requires_grad=True)
    off = (1.5 * torch.randn(B, 2 * dg * 9, H, H, device=dev)).requires_grad_(True)
    msk = torch.sigmoid(torch.randn(B, dg * 9, H, H, device=dev)).requires_grad_(True)
    gout = torch.randn(B, O, H, H, device=dev)
    ours = fwd_and_step(lambda: D.dcn_v2_conv(x, off, msk, w, b, 1, 1, 1, dg), (x, w, b, off, msk), gout)
    ref = fwd_and_step(lambda: deform_conv2d(x, off, w, b, stride=1, padding=1, dilation=1, mask=msk), (x, w, b, off, msk), gout)
    flops = 2.0 * B * H * H * O * 9 * C
    out = {"workload": "configs[2] per-GPU share: %d images; DCN 1024->512 ch 38x38 dg 4; Self_Attn core C 512 38x38 (N = M = 1444)" % B,
           "dcn": {"fwd_us": ours[0], "fwd_bwd_us": ours[1], "ref_fwd_us": ref[0], "ref_fwd_bwd_us": ref[1], "ref": "torchvision.ops.deform_conv2d fp32",
                   "fwd_tflops_algorithmic": flops / ours[0] / 1e6, "speedup_fwd": ref[0] / ours[0], "speedup_fwd_bwd": ref[1] / ours[1]}}
    Cs, N = 512, H * H
    th = (0.5 * torch.randn(B, Cs // 8, N, device=dev)).requires_grad_(True)
    ph = torch.randn(B, Cs // 8, N, device=dev, requires_grad=True)
    g = torch.randn(B, Cs // 2, N, device=dev, requires_grad=True)
    d_o = torch.randn(B, Cs // 2, N, device=dev)

    def ref_core():
        attn = torch.softmax(torch.bmm(th.permute(0, 2, 1), ph), -1)               # self_attn.py:71-72
        return torch.bmm(g, attn.permute(0, 2, 1))                                 # :80
    ours = fwd_and_step(lambda: S.attention_core(th, ph, g)[0], (th, ph, g), d_o)
    ref = fwd_and_step(ref_core, (th, ph, g), d_o)
    out["self_attn_core"] = {"fwd_us": ours[0], "fwd_bwd_us": ours[1], "ref_fwd_us": ref[0], "ref_fwd_bwd_us": ref[1],
                             "ref": "torch bmm + softmax + bmm, fp32 (cuBLAS SIMT)", "launches": "1 forward + 2 backward (reference: 5 + 8 aten kernels)",
                             "speedup_fwd": ref[0] / ours[0], "speedup_fwd_bwd": ref[1] / ours[1]}
    torch.backends.cuda.matmul.allow_tf32 = tf32
    return out


def time_model_step(a, torch, dev, B):
    """BASELINE.json configs[1] as the reference words it — a GSSD *training step*: model forward, MultiBoxLoss, backward — on a
    stand-in of the reference model (same constructors and state-dict as models/ssd_multiphase_custom_group.py, tests/gssd_standin.py;
    seeded random weights, synthetic 4-phase slices).  Two forwards of the same model object: its own torch modules (cuDNN fp32, what
    the reference runs) and `gssd_forward` (the six source chains on the tcgen05 kernels, forward AND backward); the criterion is
    ours in both.  Informational: the headline metric is the multibox head, which is 0.1 % of this step."""
    import types
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gssd_standin as G
    from grouped_ssd_pytorch_b200 import config, synthetic as syn
    from grouped_ssd_pytorch_b200.layers import MultiBoxLoss, PriorBox
    from grouped_ssd_pytorch_b200.layers.modules.source_block import gssd_forward
    torch.backends.cudnn.benchmark = True                                   # train_lesion_multiphase_v2.py:594
    net = G.StandInSSD('train', 2, True, PriorBox(config.v2).forward())
    net.load_state_dict(G.seeded_state(net.state_dict(), 71))
    net.to(dev).train()
    x = G.seeded_input(72, B).to(dev)
    targets = [torch.from_numpy(t).to(dev) for t in syn.targets(syn.rng(5), B, 1, a.gmax)]
    crit = MultiBoxLoss(2, MATCH_THRESH, True, 0, True, NEGPOS, 0.5, False, True)
    crit.process_group = False
    fast = types.MethodType(gssd_forward, net)

    def ref_forward(xx):
        loc, conf = G.forward_torch(net, xx)
        return loc, conf, net.priors

    def fast_backbone(xx):                                                  # conv3_2 .. conv5_3 on the tcgen05 kernels too (bf16, opt-in)
        return gssd_forward(net, xx, backbone=True)

    out = {}
    def measure(fwd):
        def step():
            net.zero_grad(set_to_none=True)
            o = fwd(x)
            ll, lc = crit(o, targets)
            (ll + lc).backward()
            return ll, lc
        for _ in range(3):
            ll, lc = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 8
        e0.record()
        for _ in range(n):
            ll, lc = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        return {"ms_per_step": ms, "images_per_s": B / (ms * 1e-3), "loss_l": float(ll.detach()), "loss_c": float(lc.detach())}

    for name, fwd in (("torch_forward", ref_forward), ("gssd_forward", fast)):
        out[name] = measure(fwd)
    out["speedup"] = out["torch_forward"]["ms_per_step"] / out["gssd_forward"]["ms_per_step"]
    try:                                                                    # opt-in path: its failure must not cost the two lines above
        out["gssd_forward_backbone"] = measure(fast_backbone)
        out["speedup_backbone"] = out["torch_forward"]["ms_per_step"] / out["gssd_forward_backbone"]["ms_per_step"]
    except Exception as e:                                                  # pragma: no cover
        out["gssd_forward_backbone"] = {"error": repr(e)}
    out["workload"] = "GSSD (ssd_type gssd, batch_norm, 8.34 M parameters) training step, batch %d, 4-phase 300x300: forward + MultiBoxLoss + backward, eager" % B
    return out


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
        ks = time_kernels(argparse.Namespace(mode="both"), lib, _lib, torch, dev, pri, dsets, pack_targets, B, P, iters=12)
        for k in ks:
            rows.append({"priors": pname, "batch": B, "gmax": gmax, "kernel": k["name"].split(" ")[0], "us": k["us"],
                         "gbs": k["gbs"], "frac": k["gbs"] / hbm})
        two = [k for k in ks if k["name"].startswith(("gssd_mbox_match ", "gssd_mbox_loss "))]
        if len(two) == 2:   # SURVEY §8d books matching + loss together at 64 B per prior: the fraction on that definition
            us = two[0]["us"] + two[1]["us"]
            rows.append({"priors": pname, "batch": B, "gmax": gmax, "kernel": "match+loss (64 B/prior)", "us": us,
                         "gbs": B * P * 64 / us / 1e3, "frac": B * P * 64 / us / 1e3 / hbm})
        del dsets
        torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
