"""Stage III stand-in size of the REFERENCE at num_thr=1 (its only deterministic setting) on a named bench workload,
stored in tests/golden/bits_ref_t1.json so that bench.py can print bits/base against `harc -t 1` without spending the
~10 CPU-minutes that run takes at configs[1].  Run where oracle/_ref exists:

    python tools/ref_t1_bits.py --config 1 [--seed 1]
"""
import argparse
import datetime
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
import refrun as R
import workload as W


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--reads", type=int, default=None)
    ap.add_argument("--genome", type=int, default=None)
    a = ap.parse_args()
    cfg = dict(W.CONFIGS[a.config])
    if a.reads:
        cfg["reads"] = a.reads
    if a.genome:
        cfg["genome"] = a.genome
    L = cfg["L"]
    sig = "reads=%d,L=%d,genome=%d,rc=%d,errors=%d,seed=%d" % (cfg["reads"], L, cfg["genome"], int(cfg["rc"]), int(cfg["errors"]), a.seed)
    w = W.make(cfg["reads"], L, cfg["genome"], rc=cfg["rc"], errors=cfg["errors"], seed=a.seed, keep_all=False)
    tmp = tempfile.mkdtemp(prefix="harct1")
    try:
        W.write_dir(w, tmp)
        t1, _ = R.reorder(tmp, L, 1, timeout=36000)
        t2, _ = R.encoder(tmp, L, 1, timeout=36000)
        total, parts = R.standin_size(tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    path = os.path.join(ROOT, "tests", "golden", "bits_ref_t1.json")
    fx = json.load(open(path)) if os.path.exists(path) else {}
    fx[sig] = {"standin_bytes": total, "streams": parts, "bits_per_base": 8.0 * total / (cfg["reads"] * L),
               "reorder_s": t1, "encoder_s": t2, "made": datetime.date.today().isoformat(),
               "how": "oracle/_ref L%d_T1 reorder.out + encoder.out on tools/workload.py make(%s), refrun.standin_size" % (L, sig)}
    json.dump(fx, open(path, "w"), indent=1, sort_keys=True)
    print(sig, fx[sig]["bits_per_base"], "reorder %.0fs encoder %.0fs" % (t1, t2))


if __name__ == "__main__":
    main()
