"""GPU diagnostic: run conv cases one by one (each case in its own subprocess with a timeout so a
hung kernel cannot block the rest) and print error statistics instead of asserting."""
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

CASES = {
    # name: (G, B, H, W, Cin, Cout, k, s, p, residual, relu, dtype, impl)
    "simt_f32_stem": (1, 2, 32, 32, 3, 64, 7, 2, 3, False, True, "f32", 0),
    "simt_f32_3x3_res": (1, 2, 16, 16, 64, 64, 3, 1, 1, True, True, "f32", 0),
    "simt_bf16_3x3": (1, 2, 16, 16, 64, 64, 3, 1, 1, False, True, "bf16", 1),
    "tc_1x1": (1, 2, 16, 16, 64, 64, 1, 1, 0, False, False, "bf16", 2),
    "tc_3x3": (1, 2, 16, 16, 64, 64, 3, 1, 1, False, True, "bf16", 2),
    "tc_3x3_res": (1, 2, 16, 16, 64, 64, 3, 1, 1, True, True, "bf16", 2),
    "tc_layer1": (1, 4, 64, 64, 64, 64, 3, 1, 1, True, True, "bf16", 2),
    "tc_s2": (1, 2, 32, 32, 64, 128, 3, 2, 1, False, True, "bf16", 2),
    "tc_k2_n128": (1, 2, 16, 16, 128, 128, 3, 1, 1, True, True, "bf16", 2),
    "tc_n256": (1, 3, 16, 16, 128, 256, 3, 2, 1, False, True, "bf16", 2),
    "tc_grouped_l4": (3, 2, 8, 8, 512, 512, 3, 1, 1, True, True, "bf16", 2),
    "tc_persistent": (1, 40, 16, 16, 64, 64, 3, 1, 1, False, True, "bf16", 2),
    "tc_pair": (1, 2, 32, 32, 32, 64, 3, 2, 1, False, True, "bf16", 2),
}


def run_case(name):
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import test_gpu_kernels as T
    G, B, H, W, Cin, Cout, k, s, p, residual, relu, dt, impl = CASES[name]
    dtype = torch.float32 if dt == "f32" else torch.bfloat16
    t0 = time.time()
    print(f"{name}: start", flush=True)
    y, ref = T._conv_case(G, B, H, W, Cin, Cout, k, s, p, dtype, residual, relu, bias=(Cin == 32), impl=impl)
    torch.cuda.synchronize()
    d = (y - ref).abs()
    mx = ref.abs().max().item()
    print(f"{name}: shape {tuple(y.shape)} rel_err {d.max().item() / mx:.3e}  mean_abs_err {d.mean().item():.3e} "
          f"ref_max {mx:.3f} finite {bool(torch.isfinite(y).all())}  t={time.time() - t0:.1f}s", flush=True)
    if d.max().item() / mx > 1e-2:
        bad = d > 1e-2 * mx
        print(f"   bad fraction {bad.float().mean().item():.4f}; per-image {bad.float().mean(dim=(1, 2, 3)).tolist()}")
        print(f"   bad per channel-block of 8: {bad.float().mean(dim=(0, 2, 3)).view(-1, 8).mean(1).tolist()[:16]}")
        print(f"   bad per row: {[round(v, 2) for v in bad.float().mean(dim=(0, 1, 3)).tolist()[:32]]}")
        print(f"   bad per col: {[round(v, 2) for v in bad.float().mean(dim=(0, 1, 2)).tolist()[:32]]}")
        print(f"   y[0,:4,0,0] {y[0, :4, 0, 0].tolist()} ref {ref[0, :4, 0, 0].tolist()}")
        print(f"   y[0,:4,5,5] {y[0, :4, min(5, y.shape[2] - 1), min(5, y.shape[3] - 1)].tolist()} "
              f"ref {ref[0, :4, min(5, y.shape[2] - 1), min(5, y.shape[3] - 1)].tolist()}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for name in CASES:
            try:
                r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=90)
                out = (r.stdout + ("\n" + r.stderr[-1500:] if r.returncode != 0 else "")).strip()
                print(out if out else f"{name}: no output rc={r.returncode}", flush=True)
            except subprocess.TimeoutExpired as e:
                so = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                se = e.stderr.decode() if isinstance(e.stderr, bytes) else (e.stderr or "")
                print(f"{name}: TIMEOUT (hang) stdout: {so[-300:]} stderr: {se[-600:]}", flush=True)
