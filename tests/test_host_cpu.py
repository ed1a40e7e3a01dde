"""CPU: host-side logic -- the C-ABI library loads and exports every symbol the header declares,
state-dict compatibility, flag envelope, batch sharding + all-gather over gloo (world_size 2)."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from humanliff_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "humanliff_b200.h")).read()
    declared = set(re.findall(r"\b(hl_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert lib.hl_version() >= 100
    assert lib.hl_conv_cout_pad(27) == 32 and lib.hl_conv_cout_pad(192) == 192


def test_build_recipe_covers_every_source_and_its_includes():
    """humanliff_b200/build.py: every .cu under csrc/ is compiled, and a source that #includes another source (the
    canonical-mode render kernel re-compiles render_tc5.cu) declares it in EXTRA_DEPS, so that the per-object stamp sees an
    edit of the included file."""
    from humanliff_b200 import build
    csrc = os.path.join(ROOT, "humanliff_b200", "csrc")
    on_disk = sorted(f for f in os.listdir(csrc) if f.endswith(".cu"))
    assert sorted(build.SOURCES) == on_disk, (build.SOURCES, on_disk)
    for src in on_disk:
        inc = re.findall(r'#include\s+"([^"]+\.cu)"', open(os.path.join(csrc, src)).read())
        assert sorted(inc) == sorted(build.EXTRA_DEPS.get(src, [])), (src, inc)
        for h in re.findall(r'#include\s+"([^"]+\.(?:cuh|h))"', open(os.path.join(csrc, src)).read()):
            assert os.path.exists(os.path.join(csrc, h)) or os.path.exists(os.path.join(ROOT, "include", os.path.basename(h))), (src, h)


def test_mlp_pack_offsets_match_header():
    from humanliff_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "humanliff_b200.h")).read()
    env = {}
    for m in re.finditer(r"#define (HL_MLP_[A-Z0-9_]+)\s+(.+?)\s*(?:/\*|$)", hdr, re.M):
        env[m.group(1)] = eval(m.group(2), {}, env)
    for k, v in env.items():
        assert getattr(_lib, k[3:]) == v, k


def test_production_state_dict_shape_contract():
    from humanliff_b200 import factory
    model, diffusion = factory.create_model_and_diffusion(**factory.production_flags("250"))
    sd = model.state_dict()
    assert len(sd) == 953
    assert sum(v.numel() for v in sd.values()) == 497_173_083
    assert sd["input_blocks.0.0.weight"].shape == (192, 27, 3, 3)
    assert sd["input_blocks.13.1.qkv.weight"].shape == (1152, 384, 1)
    assert sd["output_blocks.0.0.skip_connection.weight"].shape == (768, 1536, 1, 1)
    assert sd["input_blocks_proj_cond.23.weight"].shape == (768, 768, 1, 1)
    assert sd["output_blocks.3.2.conv.weight"].shape == (768, 768, 3, 3)
    assert diffusion.num_timesteps == 250 and diffusion.timestep_map[-1] == 999


def test_flag_envelope_raises_cleanly():
    from humanliff_b200 import factory
    flags = factory.production_flags()
    for bad in (dict(cond_type="concat"), dict(use_3d_aware=True), dict(use_scale_shift_norm=False)):
        with pytest.raises(NotImplementedError):
            factory.create_model_and_diffusion(**dict(flags, image_size=32, num_channels=64, **bad))
    from humanliff_b200.renderer import Renderer
    with pytest.raises(NotImplementedError):
        Renderer(triplane_ch=18, test=True)


def test_shard_batch():
    from humanliff_b200.dist import shard_batch
    spans = [shard_batch(64, r, 8) for r in range(8)]
    assert spans == [(8 * r, 8 * r + 8) for r in range(8)]
    spans = [shard_batch(10, r, 4) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from humanliff_b200.dist import all_gather_samples, shard_batch, warm_up
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
full = torch.arange(6 * 27 * 4 * 4, dtype=torch.float32).reshape(6, 27, 4, 4)
labels = torch.arange(6) - 2 + (1 << 40) * (torch.arange(6) % 2)     # negative and > 32-bit labels survive the bit-cast
warm_up((3, 27, 4, 4), "cpu")
for G in (6, 5):                                   # equal and ragged shards
    a, b = shard_batch(G, r, w)
    out, lab = all_gather_samples(full[a:b].clone(), labels[a:b].clone())
    assert torch.equal(out, full[:G]) and torch.equal(lab, labels[:G]), (G, r)
a, b = shard_batch(6, r, w)                        # caller-declared equal shards: one collective, no size exchange
for rep in range(2):                               # twice: the cached receive buffer is reused
    out, lab = all_gather_samples(full[a:b] + rep, labels[a:b], equal_shards=True)
    assert torch.equal(out, full + rep) and torch.equal(lab, labels), r
out, lab = all_gather_samples(full[a:b], None, equal_shards=True)
assert torch.equal(out, full) and lab is None
dist.barrier()
print("ok", r)
'''


def test_all_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(script), ROOT]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("ok") == 2


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("fp16", 1.5e-3)])
def test_step_plan_dataflow_by_cpu_interpretation(precision, tol):
    """The launch list `_StepPlan` compiles (pointers, pitches, concat slices, statistics rows, FiLM offsets,
    ControlNet projection wiring) interpreted on the CPU must reproduce the reference golden."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import plan_emulator
    from common import CASES, load_golden, model_state_dict, rel_l2
    from humanliff_b200.unet import _StepPlan
    fname, flags, seed, heads = CASES["tiny"]
    model, diffusion, sd = model_state_dict(dict(flags, precision=precision), seed)
    model.load_state_dict(sd)
    g = load_golden(fname)
    B, _, H, W = g["x"].shape
    cpu = torch.device("cpu")
    model._pack(cpu)
    plan = _StepPlan(model, cpu, B, H, W)
    names = [n for n, _, _ in plan.calls]
    assert names.count("hl_zero") == 1 and "hl_gn_stats" not in names
    ts = torch.tensor(diffusion.timestep_map)[torch.full((B,), 100)]
    for rep in range(2):            # twice: the statistics arena must be re-zeroed by the plan itself
        eps = plan_emulator.run_plan(plan, g["x"], ts, g["x_cond"], g["y"])
        assert rel_l2(eps, g["eps_100"]) < tol


def test_layer_handoff_files(tmp_path):
    """.npz hand-off between clothing layers (triplane_sample_layered.py:131-132,229-244): arr_0 / arr_1 keys,
    file naming, row window, lossless fp32 round trip."""
    import numpy as np
    from humanliff_b200.layered import LAYER_NAMES, layer_npz_path, load_layer_cond, save_layer_npz
    assert LAYER_NAMES == ("person", "person_pant", "person_pant_shirt", "person_pant_shirt_shoes")
    x = torch.randn(5, 27, 8, 8)
    lab = torch.full((5,), 2, dtype=torch.int64)
    p = layer_npz_path(str(tmp_path), 2, x.shape, "ema_0.9999_100000", start_id=7)
    assert p.endswith("samples_person_pant_shirt_5x27x8x8_ema_0.9999_100000_start_id_7.npz")
    save_layer_npz(p, x, lab)
    z = np.load(p)
    assert sorted(z.files) == ["arr_0", "arr_1"] and z["arr_1"].tolist() == [2] * 5
    got = load_layer_cond(p, 1, 3, "cpu")
    assert torch.equal(got, x[1:4])
    with pytest.raises(ValueError):
        load_layer_cond(p, 4, 2, "cpu")


@pytest.mark.parametrize("B", [1, 4])
def test_conv_planner_on_every_production_launch(B):
    """Host-only tiling query (hl_conv2d_plan_info) over every conv of the production step: the tcgen05 path applies
    to all of them and each plan respects the hardware limits (227 KB shared memory minus the static part, 512 TMEM
    columns, one wave of CTAs when K is split, workspace large enough for the partial sums)."""
    import ctypes
    from humanliff_b200 import _lib, factory
    from humanliff_b200.unet import _StepPlan
    model, _ = factory.create_model_and_diffusion(**factory.production_flags(""))
    model._pack(torch.device("cpu"))
    plan = _StepPlan(model, torch.device("cpu"), B, 256, 256)
    lib = _lib.load()
    n_conv = n_split = 0
    n_dual = 0
    for name, a, _br in plan.calls:
        if name == "hl_conv2d_dual":          # the 24 ControlNet projections: one launch, two results (never split)
            n_dual += 1
            Bn, H, W, Cin, Cout, k, s = a[15:22]
            has_res, want_stats, ws = 2, 1, 0       # has_res = 2: plan of the dual-output launch
        elif name == "hl_conv2d":
            Bn, H, W, Cin, Cout, k, s = a[11:18]
            has_res, want_stats, ws = int(a[5] is not None), int(a[9] is not None), plan.SPLITK_BYTES
        else:
            continue
        n_conv += 1
        out = (ctypes.c_int * 16)()
        assert lib.hl_conv2d_plan_info(1, Bn, H, W, Cin, Cout, k, s, has_res, want_stats, ws, out) == 0
        o = list(out)
        assert o[0] == 1, ("tensor-core path must apply", (Bn, H, W, Cin, Cout, k, s))
        pair, mh, n_tile, halo, a_slots, b_slots, nbuf, acc, tmem, smem, grid, tiles, S, kc = o[1:15]
        assert pair in (1, 2) and mh in (1, 2) and n_tile % 32 == 0 and 32 <= n_tile <= 256
        assert smem + (12288 if name == "hl_conv2d_dual" else 0) <= (227 - 15) * 1024     # dual: + group 1's accumulators
        assert tmem <= 512 and tmem >= acc * mh * n_tile
        assert a_slots >= 2 and b_slots >= 2 and 2 <= nbuf <= 4
        assert 1 <= grid <= 148 and tiles >= 1
        assert kc * S == Cin // 64
        if S > 1:
            n_split += 1
            Ho, Wo = H // s, W // s
            assert k == 3 and tiles <= 148 // (pair if pair == 2 else 1) * pair   # one wave
            assert S * Bn * Ho * Wo * _lib.load().hl_conv_cout_pad(Cout) * 4 <= plan.SPLITK_BYTES
            assert o[15] == 0                                    # statistics move to the second pass
    # the 1x1 skip convs of the ResBlocks whose launch fills the GPU run inside hl_gn_skip (fused with GroupNorm-1)
    n_fused = sum(1 for name, _a, _br in plan.calls if name == "hl_gn_skip")
    assert n_conv + n_fused == 256 and n_fused == (14 if B == 4 else 8) and n_dual == 24 and n_split >= 20


def test_state_dict_contract_vs_reference():
    """Checkpoint contract: the key list (in order) and every shape of the product modules'
    state_dict() equal what the reference modules produce (golden state_dict_contract.json, frozen from the
    reference by oracle/make_goldens.py contract) -- reference checkpoints load with strict=True."""
    import json
    from common import GOLDEN, PROD, TINY
    from humanliff_b200 import factory
    from humanliff_b200.renderer import ReconRenderer, Renderer
    with open(os.path.join(GOLDEN, "state_dict_contract.json")) as f:
        ref = json.load(f)

    def contract(module):
        return [[k, list(v.shape)] for k, v in module.state_dict().items()]

    # the 23 flags of script_util.model_and_diffusion_defaults (plus this package's optional `precision`)
    mine = factory.model_and_diffusion_defaults()
    for k, v in ref["model_and_diffusion_defaults"].items():
        assert k in mine and mine[k] == v, (k, v, mine.get(k))
    assert set(mine) - set(ref["model_and_diffusion_defaults"]) <= {"precision"}
    assert factory.NUM_CLASSES == ref["NUM_CLASSES"] == 4
    # space_timesteps (respace.py:7-60) incl. the ddimN form and its error cases
    from humanliff_b200 import space_timesteps
    for T, sec, want in ref["space_timesteps_cases"]:
        try:
            got = sorted(space_timesteps(T, sec))
        except Exception as e:
            got = type(e).__name__
        assert got == want, (T, sec)

    model, _ = factory.create_model_and_diffusion(**PROD)
    assert contract(model) == ref["unet_production"]
    model, _ = factory.create_model_and_diffusion(**dict(TINY, cond_type="", class_cond=False))
    assert contract(model) == ref["unet_tiny_unconditional"]
    got = dict((k, s) for k, s in contract(Renderer(use_canonical_space=False, triplane_ch=27, test=True)))
    assert got == dict((k, s) for k, s in ref["renderer_hd"]), set(got) ^ set(k for k, _ in ref["renderer_hd"])
    got = dict((k, s) for k, s in contract(ReconRenderer(use_canonical_space=False, num_instances=2, triplane_dim=256,
                                                          triplane_ch=27, test=True)))
    want = dict((k, s) for k, s in ref["renderer_rn_2_instances"])
    assert got == want, set(got) ^ set(want)


def test_reference_staging_recipe(tmp_path):
    """oracle/build_ref.py: byte-for-byte staging of the reference's hot-path packages + manifest verification.
    Needs /root/reference (this container); on a box without it the staged copy is only verified if present."""
    from oracle import build_ref
    if not os.path.isdir(os.path.join(build_ref.SRC, build_ref.PACKAGES[0])):
        if build_ref.available():
            assert build_ref.verify()
        pytest.skip("reference tree not present")
    dest = str(tmp_path / "_ref")
    man = build_ref.stage(dest=dest)
    assert man and "human_diffusion/improved_diffusion/unet.py" in man["files"]
    assert "recon_NeRF/lib/renderer.py" in man["files"] and "human_diffusion/NeRF/renderer.py" in man["files"]
    assert build_ref.verify(dest)
    with open(os.path.join(dest, "human_diffusion/improved_diffusion/nn.py"), "a") as f:
        f.write("# edited\n")
    assert not build_ref.verify(dest)            # a modified copy is detected


def test_smpl_joint_chain_matches_oracle_and_asset_is_required():
    """Host half of the canonical-space path (humanliff_b200/smpl.py): the float64 joint chain against the oracle's fp32
    restatement of get_transform_params_torch (itself pinned to the reference by the render_canon_384 golden); the
    constant block has the layout the header states; without an asset the constructor fails the way the reference's
    open() does."""
    import numpy as np
    from humanliff_b200 import _lib, synth
    from humanliff_b200.renderer import Renderer
    from humanliff_b200.smpl import SmplModel
    from oracle import render_oracle
    asset = synth.synth_smpl(5)
    m = SmplModel(asset)
    tp = synth.synth_canonical_frame(asset, 21)
    A, _ = m.joint_transforms(tp["params"]["poses"].reshape(-1).numpy(), tp["params"]["shapes"].reshape(-1).numpy())
    ref = render_oracle.joint_transforms(render_oracle.smpl_tensors(asset), tp["params"]["poses"].reshape(-1),
                                         tp["params"]["shapes"].reshape(-1))
    assert np.abs(A - ref.double().numpy()).max() < 2e-6
    c, nb, R, Th = m.frame_constants(tp["params"], tp["t_params"])
    assert c.size == _lib.smpl_consts(24) == 1018 and nb == 10 and c.dtype == np.float64
    assert np.allclose(c[:12].reshape(3, 4), A[0, :3]) and np.allclose(c[-3:], Th) and np.allclose(c[-12:-3], R.reshape(-1))
    with pytest.raises(FileNotFoundError):
        Renderer(use_canonical_space=True, triplane_ch=27, test=True, smpl_path="/nonexistent/SMPL_NEUTRAL.pkl")


def test_spatial_clusters_partition_every_vertex_exactly_once():
    """The nearest-vertex search is exact for any partition of the vertices into table slots -- provided it IS a partition."""
    import numpy as np
    from humanliff_b200 import _lib, synth
    from humanliff_b200.smpl import SmplModel, spatial_clusters
    m = SmplModel(synth.synth_smpl(5))
    sv = m.slot_vertex
    assert sv.size == m.n_clusters * m.cluster_slots and m.n_clusters == _lib.SMPL_CLUSTERS and m.cluster_slots % 4 == 0
    used = np.sort(sv[sv >= 0])
    assert np.array_equal(used, np.arange(m.n_verts))
    # degenerate inputs: fewer points than clusters, one body part only
    rs = np.random.RandomState(0)
    for n, parts in ((50, None), (1000, rs.randint(0, 3, 1000)), (777, np.zeros(777, dtype=np.int64))):
        t, slots = spatial_clusters(rs.randn(n, 3), 128, parts)
        assert t.size == 128 * slots and np.array_equal(np.sort(t[t >= 0]), np.arange(n))
