#!/usr/bin/env python
"""Generates tests/golden/cv2_frame_tail.npz: known-answer vectors of the OpenCV arithmetic behind the Frame tail
(SURVEY 8f N2): cv::undistortPoints (Frame::UndistortKeyPoints, frame.cpp:614-641) and the gemm / norm of
Frame::IsInFrustum (frame.cpp:284, 313), produced by cv2 (same library as the reference's OpenCV).

Run in the build container:  python tests/golden/make_golden_frame_tail.py
"""
import os
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(4321)
    out = {"cv2_version": np.array(cv2.__version__)}
    # --- undistortPoints(src, K, dist, R=None, P=K) on keypoint-like coordinates, several cameras
    cams = [(718.856, 718.856, 607.1928, 185.2157, [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]),       # 4 coefficients
            (517.306408, 516.469215, 318.643040, 255.313989, [0.262383, -0.953104, -0.005358, 0.002628, 1.163314]),  # TUM1, 5
            (458.654, 457.296, 367.215, 248.375, [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0, 0.01, -0.02, 0.003]),  # 8
            (400.0, 410.0, 320.0, 240.0, [0.9, -2.0, 0.01, 0.01, 5.0])]  # strong: reaches the icdist < 0 branch at the borders
    for c, (fx, fy, cx, cy, dist) in enumerate(cams):
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
        d = np.array(dist, np.float32)
        pts = np.stack([rng.uniform(-50, 1300, 3000), rng.uniform(-50, 800, 3000)], 1).astype(np.float32)
        und = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, K).reshape(-1, 2)
        out[f"und_{c}_K"], out[f"und_{c}_dist"], out[f"und_{c}_src"], out[f"und_{c}_dst"] = K, d, pts, und
    out["und_cases"] = np.array(len(cams))
    # --- gemm(Rcw, P, 1, tcw, 1) and norm(PO) on 3-vectors
    R = rng.normal(0, 1, (2000, 3, 3)).astype(np.float32)
    P = rng.normal(0, 30, (2000, 3, 1)).astype(np.float32)
    t = rng.normal(0, 5, (2000, 3, 1)).astype(np.float32)
    out["gemm_R"], out["gemm_P"], out["gemm_t"] = R, P, t
    out["gemm_out"] = np.stack([cv2.gemm(R[i], P[i], 1.0, t[i], 1.0) for i in range(len(R))])
    out["norm_out"] = np.array([cv2.norm(P[i]) for i in range(len(P))], np.float64)
    np.savez_compressed(os.path.join(HERE, "cv2_frame_tail.npz"), **out)
    print("wrote", os.path.join(HERE, "cv2_frame_tail.npz"))


if __name__ == "__main__":
    main()
