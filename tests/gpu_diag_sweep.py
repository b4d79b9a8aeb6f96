"""Not a test: per-CTA clock64 traces of the resident recurrent sweeps (DS2_TRACE_FWD/BWD), for the
store-deferral switch DS2_SWEEP_DEFER.  Prints accuracy vs the fp32 path, per-phase cycle
statistics (min / median / max over the CTAs of direction 0) and sweep times at the cfg-L layer shape."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_diag_rnn as dr  # noqa: E402

SLOTS = 16
# slot pairs (LL sweeps: 0 = operand fetch starts, 1 = first group of 4 K chunks in shared memory; barrier sweeps
# (DS2_FWD_LL=0 / DS2_BWD_LL=0): 0 = grid barrier passed, 1 = TMA issued, 9..11 = fence / red.release)
PHASES = [("prev stores(11)->fetch/barrier(0)", None), ("fetch start -> group 0 ready", (0, 1)),
          ("fetch start -> MMA starts", (0, 2)), ("mma first->last group", (2, 3)), ("last group -> commit", (3, 4)),
          ("accum wake", (4, 5)), ("tmem + st.async", (5, 6)), ("partials landed", (6, 7)), ("gates/cell", (7, 8)),
          ("8 -> 11 (barrier variants: fence + red)", (8, 11)), ("deferred fp32 stores", (11, 13))]


def trace(rnn, T, B, In, H, ncta_dir):
    x, lens, ws, bnp = dr.make(rnn, True, T, B, In, H, True, ragged=False)
    dy = torch.randn(T, B, H)
    dr.run("tf32", rnn, True, x, lens, ws, bnp, dy)
    nblk = 148
    tf = torch.zeros(nblk * T * SLOTS, dtype=torch.int64, device="cuda")
    tb = torch.zeros(nblk * T * SLOTS, dtype=torch.int64, device="cuda")
    os.environ["DS2_TRACE_FWD"] = str(tf.data_ptr())
    os.environ["DS2_TRACE_BWD"] = str(tb.data_ptr())
    dr.run("tf32", rnn, True, x, lens, ws, bnp, dy)
    del os.environ["DS2_TRACE_FWD"], os.environ["DS2_TRACE_BWD"]
    for name, tr in (("fwd", tf), ("bwd", tb)):
        a = tr.cpu().view(nblk, T, SLOTS)[:ncta_dir].double()     # direction 0
        lo, hi = 10, T - 10
        period = (a[:, lo + 1:hi, 11] - a[:, lo:hi - 1, 11]).median()
        print(f"  trace {name}: step period {int(period)} cycles (median over CTAs and steps)")
        for label, sl in PHASES:
            if sl is None:
                dlt = a[:, lo + 1:hi, 0] - a[:, lo:hi - 1, 11]
            else:
                dlt = a[:, lo:hi, sl[1]] - a[:, lo:hi, sl[0]]
            per_cta = dlt.median(1).values
            print(f"     {label:20s} min {int(per_cta.min()):6d}  med {int(per_cta.median()):6d}  max {int(per_cta.max()):6d}")
        ns = a[:, lo:hi, 12]
        skew = (ns.max(0).values - ns.min(0).values)
        order = ns.argsort(0).double()
        print(f"     arrival skew (globaltimer ns): median {skew.median():.0f}  p90 {skew.quantile(0.9):.0f};"
              f" ns resolution {(ns[0, 1:] - ns[0, :-1]).abs().min():.0f}")
        late = (ns - ns.min(0).values).median(1).values
        top = late.argsort(descending=True)[:6].tolist()
        print("     latest CTAs (median ns behind first): " + ", ".join(f"{c}:{late[c]:.0f}" for c in top), flush=True)


def main():
    print(torch.cuda.get_device_name(0))
    for ll in (1, 0):
        os.environ["DS2_FWD_LL"] = os.environ["DS2_BWD_LL"] = str(ll)
        print(f"=== DS2_FWD_LL = DS2_BWD_LL = {ll}", flush=True)
        for args in [("lstm", True, 30, 32, 128, 256, True), ("gru", True, 25, 7, 64, 256, True),
                     ("rnn", True, 19, 4, 64, 256, False), ("lstm", True, 40, 32, 512, 1024, True)]:
            try:
                dr.compare(*args)
            except Exception:
                print("[EXC]", args, traceback.format_exc(), flush=True)
        try:
            trace("lstm", 200, 32, 1024, 1024, 64)
        except Exception:
            print("[EXC] trace", traceback.format_exc(), flush=True)
        try:
            dr.timing("lstm", True, 500, 32, 1024, 1024, True)
        except Exception:
            print("[EXC] timing", traceback.format_exc(), flush=True)


if __name__ == "__main__":
    main()
