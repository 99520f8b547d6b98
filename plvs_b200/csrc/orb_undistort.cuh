// Device code only: tests/native/emu_kernels.cpp compiles the same text for the CPU (tests/native/cuda_emu.hpp).
#pragma once

// Step after extraction (§8f rank 2): Frame::UndistortKeyPoints (src/Frame.cc:1507-1553) = cv::undistortPoints(mat, mat, K, dist, Mat(), K) on
// the device-resident keypoints: OpenCV's five fixed-point iterations in double, results rounded to float (cvUndistortPointsInternal with the
// default criteria and no tilt); the oracle restatement of the same lines is pinned to the real cv2.undistortPoints.
struct UndistortParams { double fx, fy, cx, cy, k[14]; };

__global__ void __launch_bounds__(256)
k_undistort_keypoints(const plvs_keypoint* __restrict__ keys, int n, UndistortParams P, plvs_keypoint* __restrict__ keys_un)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    plvs_keypoint kp = keys[i];
    const double u = kp.x, v = kp.y;
    const double ifx = 1. / P.fx, ify = 1. / P.fy;
    double x = (u - P.cx) * ifx, y = (v - P.cy) * ify;
    const double x0 = x, y0 = y;
    const double* k = P.k;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) { x = (u - P.cx) * ifx; y = (v - P.cy) * ify; break; }
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = P.fx * x + 0 * y + P.cx, yy = 0 * x + P.fy * y + P.cy, ww = 1. / (0 * x + 0 * y + 1);
    kp.x = (float)(xx * ww); kp.y = (float)(yy * ww);
    keys_un[i] = kp;
}
