"""First-contact probe of the tcgen05 conv on a real B200: every tiling variant in its own subprocess
(a trap or hang in one variant must not poison the rest), parity vs torch on the rounded operands, then
CUDA-event timing of the dominant shape per variant.  Usage: python tools/tc_probe.py [--time]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [
    # name, (B,H,W,Cin,Cout,k,stride), mode, residual, stats, tuning(mh,n_tile,halo,epi_stats,base_off)
    ("tap1_f16", (1, 64, 64, 192, 192, 3, 1), "fp16", False, False, (1, -1, 0, -1, -1)),
    ("tap1_tf32", (1, 64, 64, 192, 192, 3, 1), "tf32", False, False, (1, -1, 0, -1, -1)),
    ("tap1_res", (1, 64, 64, 192, 192, 3, 1), "fp16", True, False, (1, -1, 0, -1, -1)),
    ("tap1_stats", (1, 64, 64, 192, 192, 3, 1), "fp16", True, True, (1, -1, 0, -1, -1)),
    ("tap2", (2, 64, 64, 192, 192, 3, 1), "fp16", True, True, (2, -1, 0, -1, -1)),
    ("tap_persist", (4, 128, 128, 192, 192, 3, 1), "fp16", True, True, (1, -1, 0, -1, -1)),
    ("stride2", (2, 64, 64, 192, 192, 3, 2), "fp16", False, True, (-1, -1, -1, -1, -1)),
    ("one_by_one", (2, 32, 32, 384, 192, 1, 1), "fp16", True, True, (-1, -1, -1, -1, -1)),
    ("cout27", (1, 64, 64, 192, 27, 3, 1), "fp16", False, False, (-1, -1, -1, -1, -1)),
    ("halo1_bo1", (1, 128, 128, 64, 64, 3, 1), "fp16", False, False, (1, -1, 1, -1, 1)),
    ("halo1_bo0", (1, 128, 128, 64, 64, 3, 1), "fp16", False, False, (1, -1, 1, -1, 0)),
    ("halo2_bo1", (2, 128, 128, 192, 192, 3, 1), "fp16", True, True, (2, 192, 1, -1, 1)),
    ("halo2_bo0", (2, 128, 128, 192, 192, 3, 1), "fp16", True, True, (2, 192, 1, -1, 0)),
    ("halo2_n96", (2, 128, 128, 192, 192, 3, 1), "fp16", True, True, (2, 96, 1, -1, -1)),
    ("halo_tf32", (1, 256, 256, 64, 64, 3, 1), "tf32", True, True, (2, -1, 1, -1, -1)),
    ("auto_256", (1, 256, 256, 192, 192, 3, 1), "fp16", True, True, (-1, -1, -1, -1, -1)),
    ("small8", (4, 8, 8, 768, 768, 3, 1), "fp16", True, True, (-1, -1, -1, -1, -1)),
    # CTA-pair MMA (cta_group::2)
    ("cta2_tap", (2, 64, 64, 192, 192, 3, 1), "fp16", False, False, (1, -1, 0, -1, -1), (-1, -1, 1)),
    ("cta2_tap_res", (2, 64, 64, 192, 192, 3, 1), "fp16", True, True, (1, -1, 0, -1, -1), (-1, -1, 1)),
    ("cta2_halo", (2, 128, 128, 192, 192, 3, 1), "fp16", True, True, (1, 192, 1, -1, -1), (-1, -1, 1)),
    ("cta2_1x1", (2, 32, 32, 384, 192, 1, 1), "fp16", True, True, (1, -1, -1, -1, -1), (-1, -1, 1)),
    ("cta2_n256", (2, 64, 64, 128, 256, 3, 1), "fp16", True, True, (1, 256, -1, -1, -1), (-1, -1, 1)),
    ("cta2_stride2", (2, 128, 128, 192, 192, 3, 2), "fp16", False, True, (1, -1, -1, -1, -1), (-1, -1, 1)),
    ("cta2_tf32", (2, 64, 64, 192, 192, 3, 1), "tf32", True, True, (1, -1, 0, -1, -1), (-1, -1, 1)),
    ("cta2_persist", (4, 256, 256, 64, 192, 3, 1), "fp16", True, True, (-1, -1, -1, -1, -1), (-1, -1, 1)),
]

TIMING = [
    # name, shape (B,H,W,Cin,Cout,k), tuning (mh,n,halo,epi_stats,base_off), tuning2 (max_stages, nbuf, cta2), residual, stats
    ("3x3 192@256 plain default", (4, 256, 256, 192, 192, 3), (-1, -1, -1, -1, -1), (-1, -1, -1), False, False),
    ("3x3 192@256 plain stages4", (4, 256, 256, 192, 192, 3), (-1, -1, -1, -1, -1), (4, -1, -1), False, False),
    ("3x3 192@256 plain stages3", (4, 256, 256, 192, 192, 3), (-1, -1, -1, -1, -1), (3, -1, -1), False, False),
    ("3x3 192@256 plain tap", (4, 256, 256, 192, 192, 3), (-1, -1, 0, -1, -1), (-1, -1, -1), False, False),
    ("3x3 192@256 res+stats default", (4, 256, 256, 192, 192, 3), (-1, -1, -1, -1, -1), (-1, -1, -1), True, True),
    ("3x3 384->192@256 stats", (4, 256, 256, 384, 192, 3), (-1, -1, -1, -1, -1), (-1, -1, -1), False, True),
    ("3x3 192@128 res+stats default", (4, 128, 128, 192, 192, 3), (-1, -1, -1, -1, -1), (-1, -1, -1), True, True),
]

def run_case(idx):
    import torch
    from test_kernels_gpu import _conv_case, _check_stats
    from common import rel_l2
    name, shape, mode, residual, stats, tuning = CASES[idx][:6]
    tuning2 = CASES[idx][6] if len(CASES[idx]) > 6 else (-1, -1, 0)
    dev = torch.device("cuda:0")
    B, H, W, Cin, Cout, k, s = shape
    y, ref, ref_r, st = _conv_case(dev, B, H, W, Cin, Cout, k, s, mode=mode, residual=residual, stats=stats, seed=7,
                                   tuning=tuning, tuning2=tuning2)
    out = {"name": name, "nan": bool(torch.isnan(y).any()), "err_rounded": rel_l2(y, ref_r), "err_fp32": rel_l2(y, ref)}
    if stats:
        try:
            _check_stats(st, y, Cout)
            out["stats"] = "ok"
        except AssertionError as e:
            out["stats"] = "BAD " + str(e)[:60]
    print("RESULT " + json.dumps(out))


def run_timing(idx):
    import torch
    from humanliff_b200 import _lib
    from humanliff_b200._lib import call
    from humanliff_b200.unet import pack_conv
    name, shape, tuning, tuning2, residual, stats = TIMING[idx]
    dev = torch.device("cuda:0")
    B, H, W, Cin, Cout, k = shape
    g = torch.Generator().manual_seed(0)
    nbytes = B * H * W * (Cin * 2 + Cout * 8)
    nbuf = max(3, min(64, int(400e6 // nbytes)))          # rotate through > 126 MB of L2
    xs = [torch.randn(B, H, W, Cin, device=dev).half() for _ in range(nbuf)]
    rs = [torch.randn(B, H, W, Cout, device=dev) for _ in range(nbuf)]
    ys = [torch.empty(B, H, W, Cout, device=dev) for _ in range(nbuf)]
    st = torch.zeros(B * Cout * 2, device=dev, dtype=torch.float64)
    prof = torch.zeros(16, device=dev, dtype=torch.int64)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    wpk, bpk = pack_conv(w, torch.zeros(Cout), Cin, "fp16", dev)
    lib = _lib.load()
    lib.hl_conv_set_tuning(*tuning)
    lib.hl_conv_set_tuning2(*tuning2)
    stream = torch.cuda.current_stream(dev)

    def launch(i):
        call("hl_conv2d", xs[i % nbuf].data_ptr(), 1, Cin, wpk.data_ptr(), bpk.data_ptr(),
             rs[i % nbuf].data_ptr() if residual else None, Cout, ys[i % nbuf].data_ptr(), Cout,
             st.data_ptr() if stats else None, Cout, B, H, W, Cin, Cout, k, 1, 0, stream.cuda_stream)
    for i in range(3):
        launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record(stream)
    for i in range(reps):
        launch(i)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * B * H * W * Cout * Cin * k * k / (ms * 1e-3) / 1e12
    gbs = (B * H * W * (Cin * 2 + Cout * 4 * (2 if residual else 1))) / (ms * 1e-3) / 1e9
    # one more launch with the in-kernel wait counters of CTA 0 switched on
    lib.hl_conv_set_profile(prof.data_ptr())
    launch(0)
    torch.cuda.synchronize()
    lib.hl_conv_set_profile(None)
    pr = prof.cpu().tolist()
    keys = ["total", "mma_wait_A", "mma_wait_tmem_empty", "mma_wait_B", "epi_wait_tmem_full", "epi_wait_res",
            "epi_barrier", "prodA_wait_empty", "prodB_wait_empty", "e0_store_drain", "tiles"]
    print("RESULT " + json.dumps({"name": name, "ms": round(ms, 4), "tflops": round(tf, 1), "hbm_gbs": round(gbs),
                                  "prof_kcycles": {k_: (round(v / 1e3, 1) if k_ != "tiles" else v) for k_, v in zip(keys, pr)}}))


def worker(kind, start):
    n = len(CASES) if kind == "case" else len(TIMING)
    for i in range(start, n):
        print("BEGIN %d" % i, flush=True)
        (run_case if kind == "case" else run_timing)(i)
        sys.stdout.flush()


def drive(kind):
    """Run all jobs of `kind` in as few worker processes as possible: a worker that traps / hangs is replaced
    and the sweep continues after the job that killed it."""
    n = len(CASES) if kind == "case" else len(TIMING)
    names = [c[0] for c in (CASES if kind == "case" else TIMING)]
    start = 0
    while start < n:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", kind, str(start)],
                               capture_output=True, text=True, timeout=150)
            out, rc = p.stdout, p.returncode
            err = p.stderr
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            err, rc = "timeout", "timeout"
        last_begin = start - 1
        for line in out.splitlines():
            if line.startswith("RESULT "):
                print(line[7:], flush=True)
            elif line.startswith("BEGIN "):
                last_begin = int(line[6:])
        done = sum(1 for l in out.splitlines() if l.startswith("RESULT "))
        if start + done >= n:
            break
        failed = start + done
        tail = [l for l in (err or "").strip().splitlines() if l.strip()][-2:]
        print(json.dumps({"name": names[failed], "FAILED": rc, "tail": tail}), flush=True)
        start = failed + 1


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--worker":
        return worker(sys.argv[2], int(sys.argv[3]))
    drive("case")
    if "--time" in sys.argv:
        drive("timing")


if __name__ == "__main__":
    main()
