"""Numpy helpers that turn a synthetic sequence (synthworld) into inputs of the C-ABI: ImageAlign feature lists and
SearchPoint candidates with ground-truth depth on the plane z = 0.  No oracle, no CUDA."""
import numpy as np
from . import abi
from .synthworld import quat_R, cam_center


def unproject(cam, px):
    v = np.array([(px[0] - cam.u0) / cam.fx, (px[1] - cam.v0) / cam.fy, 1.0])
    return v / np.linalg.norm(v)


def seed_points(cfg, xyl, T_gt, max_points=None, one_per_cell=True, margin=6):
    """Map points from FAST corners of a frame at ground-truth pose T_gt: returns dict of arrays
    px (n,2) level-0 pixels, v (n,3) bearings, pos (n,3) world, level (n,), depth (n,)."""
    cam, P = cfg["cam"], cfg["params"]
    R = quat_R(T_gt[:4])
    Cw = cam_center(T_gt)
    n = len(xyl)
    gw = int(np.ceil(cam.width / P.cell_size))
    occ = set()
    out = dict(px=[], v=[], pos=[], level=[], depth=[])
    for i in range(n):
        x, y, l = (int(t) for t in xyl[(i * 7919) % n])
        w_l, h_l = int(cam.width) >> l, int(cam.height) >> l
        if x < margin or y < margin or x >= w_l - margin or y >= h_l - margin:
            continue
        px = np.array([float(x * (1 << l)), float(y * (1 << l))])
        cell = int(px[1] / P.cell_size) * gw + int(px[0] / P.cell_size)
        if one_per_cell and cell in occ:
            continue
        v = unproject(cam, px)
        d = R.T @ v
        if abs(d[2]) < 1e-9:
            continue
        s = -Cw[2] / d[2]
        if s <= 0:
            continue
        occ.add(cell)
        out["px"].append(px); out["v"].append(v); out["pos"].append(Cw + s * d); out["level"].append(l)
        out["depth"].append(s)
        if max_points and len(out["px"]) >= max_points:
            break
    return {k: np.array(v) for k, v in out.items()}


def align_feats(pts, T_ref, invalid_every=0):
    n = len(pts["px"])
    f = np.zeros(n, abi.ALIGN_FEAT_DT)
    Cw = cam_center(np.asarray(T_ref))
    f["px"] = pts["px"]
    f["v"] = pts["v"]
    f["depth"] = np.linalg.norm(pts["pos"] - Cw, axis=1)
    f["valid"] = 1
    if invalid_every:
        f["valid"][::invalid_every] = 0
    return f


def candidates(pts, T_ref, ref_handle, fixed=True, project=True, std_frac=0.05, pred_px=None):
    n = len(pts["px"])
    c = np.zeros(n, abi.CANDIDATE_DT)
    Cw = cam_center(np.asarray(T_ref))
    depth = np.linalg.norm(pts["pos"] - Cw, axis=1)
    c["ref_frame"] = ref_handle
    c["ref_T"] = np.asarray(T_ref)
    c["ref_px"] = pts["px"]
    c["ref_v"] = pts["v"]
    c["idepth"] = 1.0 / depth
    c["idepth_std"] = std_frac / depth
    c["pos"] = pts["pos"]
    c["ref_level"] = pts["level"]
    c["flags"] = (abi.CAND_FIXED if fixed else 0) | (abi.CAND_PROJECT if project else 0)
    if pred_px is not None:
        c["px"] = pred_px
    return c
