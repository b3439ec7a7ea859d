// Arithmetic that must round exactly like the scalar CPU code (translation units including this are built -fmad=false).
#pragma once
// Eigen QuaternionBase::_transformVector in fp64: uv = 2 (u x v); v + w uv + u x uv  (same association as the CPU code)
__device__ __forceinline__ void qrot_d(const double q[4] /*w,x,y,z*/, double vx, double vy, double vz, double& ox, double& oy, double& oz) {
  double ux = q[1], uy = q[2], uz = q[3], w = q[0];
  double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
  cx = cx + cx; cy = cy + cy; cz = cz + cz;
  double dx = uy * cz - uz * cy, dy = uz * cx - ux * cz, dz = ux * cy - uy * cx;
  ox = (vx + cx * w) + dx; oy = (vy + cy * w) + dy; oz = (vz + cz * w) + dz;
}

// Point_cloud_registration::refine_blur (point_cloud_registration.hpp:128-141) with m_if_motion_deblur = 1: float arithmetic (the time stamps
// are converted to float at the call), not clamped below 0, 1.0 when not finite or > 1.
__device__ __forceinline__ float refine_blur_f(float in_blur, float min_blur, float max_blur) {
  const float res = (in_blur - min_blur) / (max_blur - min_blur);
  if (!isfinite(res) || (double)res > 1.0) return 1.0f;
  return res;
}
