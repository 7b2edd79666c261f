"""CPU oracle for the pose stage — TEST INFRASTRUCTURE ONLY (see oracle/oracle.py's header).

The reference's pose solver is a third-party call: ``cv2.solvePnPRansac`` (OpenCV, present in this
image: cv2 4.13) with ``flags=cv2.SOLVEPNP_EPNP, iterationsCount=10000`` —
src/utils/metric_utils.py:121-204.  ``ransac_pnp`` below follows those lines literally;
``refined`` then minimises the reprojection error over the RANSAC inliers
(``cv2.solvePnPRefineLM``), which is the well-defined quantity a different RANSAC implementation
can be compared against: sampling is random in both, the inlier set and the least-squares pose on
it are not."""
import cv2
import numpy as np


def ransac_pnp(K, pts_2d, pts_3d, scale=1, pnp_reprojection_error=5):
    """metric_utils.py:169-204 (the OpenCV branch): returns (pose [3,4], pose_homo [4,4], inliers, state)."""
    dist_coeffs = np.zeros(shape=[8, 1], dtype="float64")
    pts_2d = np.ascontiguousarray(pts_2d.astype(np.float64))
    pts_3d = np.ascontiguousarray(pts_3d.astype(np.float64))
    K = K.astype(np.float64)
    pts_3d = pts_3d * scale
    try:
        _, rvec, tvec, inliers = cv2.solvePnPRansac(pts_3d, pts_2d, K, dist_coeffs,
                                                    reprojectionError=pnp_reprojection_error,
                                                    iterationsCount=10000, flags=cv2.SOLVEPNP_EPNP)
        rotation = cv2.Rodrigues(rvec)[0]
        tvec = tvec / scale
        pose = np.concatenate([rotation, tvec], axis=-1)
        pose_homo = np.concatenate([pose, np.array([[0, 0, 0, 1]])], axis=0)
        if inliers is None:
            inliers = np.array([]).astype(bool)
        return pose, pose_homo, inliers, True
    except cv2.error:
        return np.eye(4)[:3], np.eye(4), np.array([]).astype(bool), False


def refined(K, pts_2d, pts_3d, pose, inliers, scale=1):
    """Least-squares pose on the inlier set (LM on the reprojection error), same units as `pose`."""
    idx = np.asarray(inliers).reshape(-1)
    p3 = np.ascontiguousarray(pts_3d[idx].astype(np.float64)) * scale
    p2 = np.ascontiguousarray(pts_2d[idx].astype(np.float64))
    rvec = cv2.Rodrigues(pose[:, :3].astype(np.float64))[0]
    tvec = (pose[:, 3:].astype(np.float64) * scale).copy()
    rvec, tvec = cv2.solvePnPRefineLM(p3, p2, K.astype(np.float64), np.zeros((8, 1)), rvec, tvec,
                                      criteria=(cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, 200, 1e-12))
    return np.concatenate([cv2.Rodrigues(rvec)[0], tvec / scale], axis=-1)


def synthetic_frames(batch, n_matches=300, outlier_frac=0.3, noise_px=0.5, seed=0, hw=(512, 512)):
    """Planted pose workload: per frame a random pose, `n_matches` 3D points projected with pixel
    noise, `outlier_frac` of them replaced by uniformly random pixels.  Returns the matcher-style
    lists (m_bids, mkpts_3d_db, mkpts_query_f), intrinsics [B,3,3] and the ground-truth poses."""
    rng = np.random.default_rng(seed)
    h, w = hw
    K = np.array([[600.0, 0, w / 2], [0, 600.0, h / 2], [0, 0, 1]])
    bids, p3s, p2s, poses = [], [], [], []
    for b in range(batch):
        n = n_matches + 17 * b
        rvec = rng.normal(size=3) * 0.6
        R = cv2.Rodrigues(rvec)[0]
        t = np.array([rng.normal() * 0.05, rng.normal() * 0.05, 1.2 + 0.3 * rng.random()])
        P = (rng.random((n, 3)) - 0.5) * 0.4
        X = P @ R.T + t
        uv = (X[:, :2] / X[:, 2:]) * np.array([K[0, 0], K[1, 1]]) + K[:2, 2]
        uv += rng.normal(size=uv.shape) * noise_px
        out = rng.random(n) < outlier_frac
        uv[out] = rng.random((int(out.sum()), 2)) * np.array([w, h])
        bids.append(np.full(n, b)), p3s.append(P), p2s.append(uv)
        poses.append(np.concatenate([R, t[:, None]], 1))
    return (np.concatenate(bids).astype(np.int64), np.concatenate(p3s).astype(np.float32),
            np.concatenate(p2s).astype(np.float32), np.stack([K] * batch).astype(np.float32), np.stack(poses))
