"""pytest -m gpu: the on-device RANSAC-PnP front end (opp_pnp_ransac through
onepose_plus_plus_b200.pnp) against the reference's solver — cv2.solvePnPRansac(EPnP) as called in
src/utils/metric_utils.py:121-204 — on planted poses with 30 % outliers.  RANSAC sampling is
random in both; what must agree is the inlier set and the least-squares pose on it (tolerance 1e-3
on rotation entries and relative translation, as for the matcher)."""
import numpy as np
import pytest
import torch

from onepose_plus_plus_b200 import pnp
from oracle import pnp as opnp

pytestmark = pytest.mark.gpu


def _frames(batch, **kw):
    b, p3, p2, K, gt = opnp.synthetic_frames(batch, **kw)
    dev = torch.device("cuda")
    return (b, p3, p2, K, gt), [torch.as_tensor(x, device=dev) for x in (b, p3, p2, K)]


@pytest.mark.parametrize("scale", [1.0, 1000.0])
def test_planted_poses_match_cv2(scale):
    (b, p3, p2, K, gt), (tb, t3, t2, tK) = _frames(6, outlier_frac=0.3, noise_px=0.5, seed=3)
    p3s = p3 / scale      # the bank is stored in other units; `scale` brings it back (metric_utils.py:179,193)
    r = pnp.ransac_pnp_batched(tb, t3 / scale, t2, tK, scale=scale, reprojection_error=5.0)
    torch.cuda.synchronize()
    pose = r["pose"].double().cpu().numpy()
    mask = r["inlier_mask"].cpu().numpy()
    assert r["state"].all().item()
    for i in range(6):
        m = b == i
        ref_pose, _, inl, ok = opnp.ransac_pnp(K[i], p2[m], p3s[m], scale=scale, pnp_reprojection_error=5)
        assert ok
        ref_set = set(np.asarray(inl).reshape(-1).tolist())
        got_set = set(np.nonzero(mask[m])[0].tolist())
        assert len(ref_set ^ got_set) <= max(2, len(ref_set) // 100), (i, len(ref_set), len(got_set))
        assert int(r["n_inliers"][i].item()) == len(got_set)
        ref = opnp.refined(K[i], p2[m], p3s[m], ref_pose, inl, scale=scale)
        err_R = np.abs(pose[i][:, :3] - ref[:, :3]).max()
        err_t = np.linalg.norm(pose[i][:, 3] - ref[:, 3]) / np.linalg.norm(ref[:, 3])
        print(f"frame {i}: inliers {len(got_set)}/{int(m.sum())}  |dR| {err_R:.2e}  |dt|/|t| {err_t:.2e}")
        assert err_R <= 1e-3 and err_t <= 1e-3
        # and both sit at the planted pose up to the pixel noise
        gt_i = gt[i].copy()
        gt_i[:, 3] /= scale
        assert np.abs(pose[i][:, :3] - gt_i[:, :3]).max() < 5e-3


def test_degenerate_frames_and_determinism():
    (b, p3, p2, K, gt), (tb, t3, t2, tK) = _frames(4, n_matches=120, seed=5)
    # frame 1 keeps only 3 matches (below the minimal sample), frame 2 none at all
    keep = np.ones(len(b), dtype=bool)
    idx1 = np.nonzero(b == 1)[0]
    keep[idx1[3:]] = False
    keep[b == 2] = False
    sel = torch.as_tensor(keep, device="cuda")
    args = (tb[sel], t3[sel], t2[sel], tK)
    r1 = pnp.ransac_pnp_batched(*args, seed=7)
    r2 = pnp.ransac_pnp_batched(*args, seed=7)
    torch.cuda.synchronize()
    assert r1["state"].cpu().tolist() == [True, False, False, True]
    eye = torch.eye(4, device="cuda")[:3]
    assert torch.equal(r1["pose"][1], eye) and torch.equal(r1["pose"][2], eye)
    assert r1["n_inliers"].cpu().tolist()[1:3] == [0, 0]
    for k in ("pose", "n_inliers", "inlier_mask"):
        assert torch.equal(r1[k], r2[k]), f"{k} must not depend on scheduling"
    # an empty match list (M = 0) does not raise
    e = pnp.ransac_pnp_batched(tb[:0], t3[:0], t2[:0], tK)
    assert not e["state"].any().item() and e["inlier_mask"].numel() == 0


def test_reference_signatures():
    """ransac_PnP / compute_query_pose_errors keep the reference's call shapes (metric_utils.py:121,207)."""
    (b, p3, p2, K, gt), (tb, t3, t2, tK) = _frames(3, seed=9)
    m = b == 1
    pose, pose_homo, inliers, state = pnp.ransac_PnP(K[1], p2[m], p3[m], scale=1, pnp_reprojection_error=5,
                                                     img_hw=[512, 512], use_pycolmap_ransac=False)
    assert state and pose.shape == (3, 4) and pose_homo.shape == (4, 4) and inliers.ndim == 2
    assert np.abs(pose - gt[1]).max() < 5e-3
    gt_h = np.tile(np.eye(4), (3, 1, 1))
    gt_h[:, :3] = gt
    data = {"m_bids": tb, "mkpts_3d_db": t3, "mkpts_query_f": t2, "query_intrinsic": tK,
            "query_pose_gt": torch.as_tensor(gt_h), "q_hw_i": torch.Size((512, 512)),
            "query_image_scale": torch.ones(3, 2)}
    pnp.compute_query_pose_errors(data, {"pnp_reprojection_error": 5, "point_cloud_rescale": 1,
                                         "use_pycolmap_ransac": False})
    assert len(data["R_errs"]) == 3 and max(data["R_errs"]) < 0.5 and max(data["t_errs"]) < 0.5   # deg, cm
    assert data["pose_pred"].shape == (3, 4, 4) and len(data["inliers"][0]) > 100
