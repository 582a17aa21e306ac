"""Per-phase clock64 breakdown of one CTA of each kernel (needs the --phase-timing debug build)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grouped_ssd_pytorch_b200 import build
build.LIB = build.LIB.replace(".so", "_dbg.so")
build.stale = lambda: False
import torch
from grouped_ssd_pytorch_b200 import _lib
lib = _lib.require_cuda()
import runpy
for args in (["32", "v2", "5", "2"], ["256", "v2", "5", "2"], ["1024", "v2", "5", "2"], ["64", "v2_512", "32", "2"]):
    sys.argv = ["ncu_target.py"] + args
    runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_target.py"), run_name="__main__")
    for name, labels in (("fused", ["conf->smem+max", "IoU sweep", "exchange+force+count", "wait xmax+keys", "select", "wait N+sweep2", "finish"]),
                         ("match", ["sweep", "reduce+force", "emit", "tail-sync"]),
                         ("loss", ["sweep1", "select", "sweep2", "finish"]),
                         ("detect", ["threshold", "select", "collect", "sort", "decode", "mask", "resolve", "emit"])):
        buf = (ctypes.c_longlong * 32)()
        getattr(lib, "gssd_debug_phase_clocks_" + name)(buf)
        t = list(buf)
        d = [t[i + 1] - t[i] for i in range(len(labels))]
        print("B=%s %s %-7s total %7d cyc: " % (args[0], args[1], name, t[len(labels)] - t[0]) +
              "  ".join("%s=%d" % (l, x) for l, x in zip(labels, d)), flush=True)
        if name == "fused" and hasattr(lib, "gssd_debug_fused_cta_ns"):
            import numpy as np
            ns = (ctypes.c_ulonglong * 2048)()
            lib.gssd_debug_fused_cta_ns(ns)
            arr = np.array(list(ns), dtype=np.int64).reshape(2, 1024)
            n = int((arr[0] > 0).sum())
            st, en = arr[0, :n] - arr[0, :n].min(), arr[1, :n] - arr[0, :n].min()
            print("      %d CTAs, ns after the first CTA's entry: entry median %d / max %d; exit min %d / median %d / max %d; own work (exit - entry, CTA 0 excluded) median %d / max %d" % (
                n, np.median(st), st.max(), en.min(), np.median(en), en.max(), np.median((en - st)[1:]), (en - st)[1:].max()), flush=True)
        if name == "fused":
            print("      A: issue+GT=%d cp.async wait+sync=%d cluster_arrive=%d max+publish=%d | select: exchange1=%d scan=%d local gather=%d copy+sync+flatten=%d rank=%d" % (
                t[12] - t[0], t[13] - t[12], t[14] - t[13], t[1] - t[14], t[8] - t[4], t[9] - t[8], t[10] - t[9], t[11] - t[10], t[5] - t[11]), flush=True)
        if name == "loss":     # stamps inside the select: [zeroed, counted, pushed+barrier, suffix] per pass, then [gathered+barrier, ranked]
            sel = [x for x in t[8:24] if x > 0]
            print("      select stamps (cycles after sweep 1): " + " ".join(str(x - t[1]) for x in sel), flush=True)
