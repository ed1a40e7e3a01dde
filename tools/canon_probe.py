"""Canonical-space rendering (use_canonical_space=True): parity against the reference golden and throughput of a whole
256 x 256 frame (65,536 rays x 256 samples, each sample snapped to its nearest of 6,890 body vertices).
    python tools/canon_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from common import load_golden, rel_l2  # noqa: E402
from humanliff_b200 import synth  # noqa: E402
from humanliff_b200.renderer import Renderer, render  # noqa: E402

dev = torch.device("cuda:0")
gz = load_golden("render_canon_384.npz")
asset = synth.synth_smpl(int(gz["seed_smpl"]))
tp = synth.synth_canonical_frame(asset, int(gz["seed_pose"]))
mv = lambda v: {k: mv(x) for k, x in v.items()} if isinstance(v, dict) else v.to(dev)
tpd = mv(tp)
planes = synth.synth_triplane(256, seed=7).to(dev)
n = int(gz["n_rays"])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for precision in ("fp16", "fp32"):
    r = Renderer(use_canonical_space=True, triplane_ch=27, test=True, smpl=asset, precision=precision)
    shapes = {k: v.shape for k, v in r.state_dict().items() if not k.startswith("view_enc")}
    r.load_state_dict(synth.synth_state_dict(shapes, seed=int(gz["seed_w"]), weight_gain=1.5), strict=False)
    r.to(dev)
    ro, rd, near, far, u = synth.synth_canonical_rays(tp, n)
    out = render(rays_o=ro[None].to(dev), rays_d=rd[None].to(dev), near=near[None].to(dev), far=far[None].to(dev),
                 tri_planes=planes, tp_input=tpd, renderer=r, n_samples=128, n_importance=128, u=u.to(dev))
    print("[%s] parity vs the unmodified reference (384 rays): rgb %.2e  acc %.2e  depth %.2e" %
          (precision, rel_l2(out[0][0], gz["rgb"]), rel_l2(out[1][0], gz["acc"]), rel_l2(out[3][0], gz["depth"])))
    for HW in (256, 512):
        wb = tp["world_bounds"][0].tolist()
        ro, rd, near, far, hit = synth.synth_camera_rays(HW, HW, focal=300.0 * HW / 256, azimuth_deg=30.0, bounds=wb)
        args = dict(rays_o=ro[None].to(dev), rays_d=rd[None].to(dev), near=near[None].to(dev), far=far[None].to(dev),
                    tri_planes=planes, tp_input=tpd, renderer=r, n_samples=128, n_importance=128)
        ms = timed(lambda: render(**args))
        r2 = Renderer(use_canonical_space=False, triplane_ch=27, test=True, precision=precision)
        r2.load_state_dict(r.state_dict(), strict=False)
        r2.to(dev)
        args2 = dict(args, renderer=r2, tp_input={"world_bounds": tpd["world_bounds"]})
        ms2 = timed(lambda: render(**args2))
        print("[%s] %d x %d frame (%d rays, %.0f%% hit the box): canonical space %.1f ms = %.2f M rays/s; the same rays "
              "without the deformation %.1f ms = %.2f M rays/s" % (precision, HW, HW, ro.shape[0],
              100.0 * float(hit.float().mean()), ms, ro.shape[0] / ms / 1e3, ms2, ro.shape[0] / ms2 / 1e3))
ms = timed(lambda: r.smpl.frame_tables(tpd, 0, dev), reps=10)
print("per-frame vertex tables (host joint chain + hl_smpl_vertex_tables): %.2f ms" % ms)

# phase counters of the tcgen05 kernel in canonical mode (cycles of CTA 0 / group 0; "gather" includes the deformation)
from humanliff_b200 import _lib  # noqa: E402
r = Renderer(use_canonical_space=True, triplane_ch=27, test=True, smpl=asset, precision="fp16")
r.load_state_dict(synth.synth_state_dict(shapes, seed=int(gz["seed_w"]), weight_gain=1.5), strict=False)
r.to(dev)
wb = tp["world_bounds"][0].tolist()
ro, rd, near, far, hit = synth.synth_camera_rays(512, 512, focal=600.0, azimuth_deg=30.0, bounds=wb)
args = dict(rays_o=ro[None].to(dev), rays_d=rd[None].to(dev), near=near[None].to(dev), far=far[None].to(dev),
            tri_planes=planes, tp_input=tpd, renderer=r, n_samples=128, n_importance=128)
render(**args)
prof = torch.zeros(16, device=dev, dtype=torch.int64)
lib = _lib.load()
lib.hl_render5_set_profile(prof.data_ptr())
render(**args)
torch.cuda.synchronize()
lib.hl_render5_set_profile(None)
p = prof.cpu().tolist()
rays = (262144 // 148 + 1) // 2
names = ["setup", "gather+deform", "mlp", "resample+sort", "composite", "total", "mlp:barrier+issue", "mlp:wait_mma"]
print({k: round(v / rays) for k, v in zip(names, p)}, "cycles per ray (CTA 0, group 0), canonical mode")
calls = max(p[15], 1)
print("nearest-vertex search per call (thread 0): pass 1 %d, phase A %d, candidate mask %d, phase B %d cycles; clusters in "
      "phase A %.2f, candidates %.2f, scanned %.2f; %d calls" % (p[8] / calls, p[9] / calls, p[10] / calls, p[11] / calls,
                                                                 p[12] / calls, p[13] / calls, p[14] / calls, calls))
