"""SMPL body model tables for the canonical-space renderer (``use_canonical_space=True``).

Replaces, for the inference path, ``read_pickle`` / ``SMPL_to_tensor`` / ``get_transform_params_torch`` /
``get_rigid_transformation_torch`` / ``batch_rodrigues_torch`` of human_diffusion/NeRF/renderer.py:333-433 (identical in
recon_NeRF/lib/renderer.py:352-433) and the per-frame part of ``deform_target2c_op`` (:52-113).

Split of the work: the kinematic chain of the J joints (24 for SMPL) is a sequential product of 4x4 matrices -- it runs
here on the host in float64, twice per frame (the frame's pose; the canonical "big pose" with zero shape), and is
uploaded as one small constant block.  Everything per *vertex* (skinning-weight blends of the joint transforms, the pose /
shape blend-shape offsets, the inverse, the composition into one 3x4 affine) and everything per *point* (nearest vertex,
the affine) is CUDA: ``hl_smpl_vertex_tables`` and ``hl_render_rays_canon`` / ``hl_canonical_points``.
"""
import ctypes
import os
import pickle

import numpy as np
import torch

from . import _lib
from ._lib import call


def read_asset(path):
    """``assets/SMPL_NEUTRAL.pkl`` (latin1 pickle, renderer.py:333-337) or an SMPL-X ``.npz`` (:46-48) -> dict."""
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path}: use_canonical_space=True needs the SMPL body model the reference loads from this path "
            "(human_diffusion/NeRF/renderer.py:41-50); pass smpl=<dict of its arrays> or smpl_path=...")
    if path.endswith(".npz"):
        return dict(np.load(path, allow_pickle=True))
    with open(path, "rb") as f:
        u = pickle._Unpickler(f)
        u.encoding = "latin1"
        return u.load()


def _dense(a):
    return np.asarray(a.toarray() if hasattr(a, "toarray") else a, dtype=np.float64)


def axis_angle_matrices(rv):
    """[N,3] float64 -> [N,3,3]; batch_rodrigues(_torch): angle = |rv + 1e-8|, R = I + sin K + (1 - cos) K^2."""
    rv = np.asarray(rv, dtype=np.float64).reshape(-1, 3)
    ang = np.linalg.norm(rv + 1e-8, axis=1, keepdims=True)
    k = rv / ang
    K = np.zeros((rv.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    return np.eye(3)[None] + np.sin(ang)[:, :, None] * K + (1.0 - np.cos(ang))[:, :, None] * (K @ K)


def spatial_clusters(points, n_clusters, part=None):
    """Partition the template vertices into ``n_clusters`` compact groups of near-equal size: vertices are first grouped by
    body part (``part`` [V]: the joint with the largest skinning weight -- a part moves almost rigidly, whereas template
    neighbours of different parts, e.g. the two thighs, separate under a pose), every part gets clusters in proportion to
    its size, and inside a part a k-d split cuts the largest extent at the matching quantile.
    -> (slot_vertex int32 [n_clusters * slots], slots): the vertex of every table slot in cluster order, -1 = unused.
    The nearest-vertex search is exact for ANY partition; compact clusters make its bounding-sphere pruning effective."""
    n = points.shape[0]
    part = np.zeros(n, dtype=np.int64) if part is None else np.asarray(part)
    groups = [np.nonzero(part == g)[0] for g in np.unique(part)]
    # largest-remainder allocation, at least one cluster per part
    share = np.array([g.size for g in groups], dtype=np.float64) * n_clusters / n
    k = np.maximum(1, np.floor(share).astype(int))
    while k.sum() > n_clusters:
        k[np.argmax(k - share)] -= 1
    while k.sum() < n_clusters:
        k[np.argmax(share - k)] += 1
    leaves = []

    def split(idx, kk):
        if kk == 1 or idx.size <= 1:
            leaves.append(idx)
            for _ in range(kk - 1):
                leaves.append(idx[:0])
            return
        p = points[idx]
        order = idx[np.argsort(p[:, int(np.argmax(p.max(0) - p.min(0)))], kind="stable")]
        k1 = kk // 2
        cut = int(round(order.size * k1 / kk))
        split(order[:cut], k1)
        split(order[cut:], kk - k1)

    for g, kk in zip(groups, k):
        split(g, int(kk))
    slots = max(4, (max(l.size for l in leaves) + 3) // 4 * 4)
    table = np.full((n_clusters, slots), -1, dtype=np.int32)
    for c, idx in enumerate(leaves):
        table[c, : idx.size] = np.sort(idx)
    return table.reshape(-1), slots


class SmplModel:
    """The asset's tables: fp32 on the device for the per-vertex kernel, float64 on the host for the joint chain."""

    def __init__(self, params):
        p = params
        self.v_template = _dense(p["v_template"])                       # [V,3]
        self.n_verts = self.v_template.shape[0]
        sd = _dense(p["shapedirs"])
        self.shapedirs = sd.reshape(self.n_verts, 3, -1)                 # [V,3,S]
        self.weights = _dense(p["weights"])                             # [V,J]
        self.n_joints = self.weights.shape[1]
        self.posedirs = _dense(p["posedirs"]).reshape(self.n_verts, 3, -1)
        if self.posedirs.shape[2] != 9 * (self.n_joints - 1):
            raise ValueError(f"posedirs has {self.posedirs.shape[2]} pose features, expected {9 * (self.n_joints - 1)}")
        jr = _dense(p["J_regressor"])                                   # [J,V]
        self.parents = [int(v) for v in np.asarray(p["kintree_table"])[0][: self.n_joints]]
        # joints = J_regressor (v_template + shapedirs betas) is linear in betas: regress the tables once
        self.j_template = jr @ self.v_template                          # [J,3]
        self.j_shapedirs = np.einsum("jv,vcs->jcs", jr, self.shapedirs)  # [J,3,S]
        self.faces = np.asarray(p["f"]).astype(np.int64) if "f" in p else None
        self.n_clusters = _lib.SMPL_CLUSTERS
        self.slot_vertex, self.cluster_slots = spatial_clusters(self.v_template, self.n_clusters,
                                                                 np.argmax(self.weights, axis=1))
        self._dev = {}

    def device_tables(self, device):
        key = str(device)
        if key not in self._dev:
            f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
            self._dev[key] = (f(self.weights), f(self.posedirs), f(self.shapedirs),
                              torch.from_numpy(self.slot_vertex).to(device))
        return self._dev[key]

    def joint_transforms(self, poses, betas):
        """get_transform_params_torch + get_rigid_transformation_torch (renderer.py:354-420) for one sample:
        -> (A [J,4,4] float64: posed joint frames with the rest joints removed, rot_mats [J,3,3])."""
        betas = np.asarray(betas, dtype=np.float64).reshape(-1)
        J = self.j_template + self.j_shapedirs[..., :betas.size] @ betas
        rot = axis_angle_matrices(np.asarray(poses, dtype=np.float64).reshape(-1, 3)[: self.n_joints])
        G = np.zeros((self.n_joints, 4, 4))
        for j in range(self.n_joints):
            T = np.eye(4)
            T[:3, :3] = rot[j]
            T[:3, 3] = J[j] - (J[self.parents[j]] if j else 0.0)
            G[j] = T if j == 0 else G[self.parents[j]] @ T
        G[:, :3, 3] -= np.einsum("jab,jb->ja", G[:, :3, :3], J)
        return G, rot

    def frame_constants(self, params, t_params, b=0):
        """The fp64 constant block of ``hl_smpl_vertex_tables`` (HL_SMPL_CONSTS(J)) for sample ``b`` of ``tp_input``."""
        cpu = lambda t: np.asarray(torch.as_tensor(t).detach().cpu().double().numpy())
        poses, betas = cpu(params["poses"])[b].reshape(-1), cpu(params["shapes"])[b].reshape(-1)
        tposes = cpu(t_params["poses"])[b].reshape(-1)
        if betas.size > 16:
            raise NotImplementedError("more than 16 shape coefficients")
        A, rot = self.joint_transforms(poses, betas)
        Ab, trot = self.joint_transforms(tposes, np.zeros_like(betas))    # big pose with the mean shape (:95-96)
        eye = np.eye(3)[None]
        R, Th = cpu(params["R"])[b].reshape(3, 3), cpu(params["Th"])[b].reshape(3)
        bpad = np.zeros(16)
        bpad[:betas.size] = betas
        c = np.concatenate([A[:, :3, :].reshape(-1), Ab[:, :3, :].reshape(-1), (rot[1:] - eye).reshape(-1),
                            (trot[1:] - eye).reshape(-1), bpad, R.reshape(-1), Th])
        assert c.size == _lib.smpl_consts(self.n_joints)
        return c, betas.size, R.astype(np.float32), Th.astype(np.float32)

    @torch.no_grad()
    def frame_tables(self, tp_input, b, device):
        """-> dict(knn, aff, n_verts, rot, trans): the device tables of one frame + the host constants of the launch."""
        c, n_betas, R, Th = self.frame_constants(tp_input["params"], tp_input["t_params"], b)
        w, pd, sd, slots = self.device_tables(device)
        verts = torch.as_tensor(tp_input["vertices"])[b].detach().to(device, torch.float32).contiguous()
        if verts.shape != (self.n_verts, 3):
            raise ValueError(f"tp_input['vertices'] is {tuple(verts.shape)}, the asset has {self.n_verts} vertices")
        consts = torch.from_numpy(c).to(device)
        nc, cl = self.n_clusters, self.cluster_slots
        knn = torch.empty((nc + nc * cl) * 4, device=device, dtype=torch.float32)
        aff = torch.empty(self.n_verts * 12, device=device, dtype=torch.float32)
        call("hl_smpl_vertex_tables", w.data_ptr(), pd.data_ptr(), sd.data_ptr(), self.shapedirs.shape[2], n_betas,
             verts.data_ptr(), consts.data_ptr(), self.n_verts, self.n_joints, slots.data_ptr(), nc, cl, knn.data_ptr(),
             aff.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        rot = (ctypes.c_float * 9)(*[float(v) for v in R.reshape(-1)])
        trans = (ctypes.c_float * 3)(*[float(v) for v in Th])
        return {"knn": knn, "aff": aff, "n_clusters": nc, "cluster_slots": cl, "rot": rot, "trans": trans,
                "_keep": (consts, verts)}

    def table_args(self, canon):
        """The (knn_table, affine_table, n_clusters, cluster_slots, rot, trans) run of the C-ABI calls."""
        return (canon["knn"].data_ptr(), canon["aff"].data_ptr(), canon["n_clusters"], canon["cluster_slots"],
                ctypes.cast(canon["rot"], ctypes.c_void_p), ctypes.cast(canon["trans"], ctypes.c_void_p))
