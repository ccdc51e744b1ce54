// slam_wrapper — the hot-path subset of the reference's pybind11 module of the same name
// (slam/src/slam_wrapper.cpp:192-324), built on liblsdreg's C ABI (include/lsdreg.h).
//
//   pointcloud_align(source_point f32[N,4], target_point f32[M,4], guess f32[4,4]) -> f32[4,4]
//       slam_wrapper.cpp:241-242 / graph_utils.cpp:20-46; the one matcher entry Python calls directly
//       (map_manager.py:190-192 `keyframe_align`).  Same argument names, same 50 m guess guard, same
//       solver settings (eps 1e-2, 64 iterations, 20 neighbours, PCL GICP's 5 m correspondence gate);
//       the GICP cost is minimised by fast_gicp's LM instead of PCL's BFGS.
//   registration_align(method, source, target, guess, max_corr, max_process_time_us)
//       -> (T f32[4,4], converged, fitness): what every C++ caller does with the object handed out by
//       select_registration_method (registrations.hpp:15-16): setInputTarget/Source, align,
//       hasConverged, getFinalTransformation, getFitnessScore — exposed for Python tests and tools.
//   dump_keyframe(directory, stamp, id, points_input f32[N,4], pose_input f32[4,4])
//       slam_wrapper.cpp:273-276 / graph_utils.cpp:123-131 -> KeyFrame::save: cloud.pcd + data in `directory`
//       (map_manager.py:286-288 calls it once per key frame when a map is saved).
// Everything else in the reference module (init_slam / process / graph editing / export) drives
// subsystems that are out of scope and is not provided here.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <cmath>
#include <stdexcept>
#include <string>

#include "lsdreg.h"

namespace py = pybind11;
using farray = py::array_t<float, py::array::c_style | py::array::forcecast>;

namespace {

struct Reg {
  lsd_reg_t* h = nullptr;
  Reg(int kind, const lsd_reg_params_t& p) { (void)kind; if (lsd_reg_create(&h, &p) != LSD_OK) throw std::runtime_error(lsd_last_error()); }
  ~Reg() { lsd_reg_destroy(h); }
};

void check_cloud(const farray& a, const char* name) {
  if (a.ndim() != 2 || a.shape(1) != 4 || a.shape(0) < 1) throw std::invalid_argument(std::string(name) + " must be float32 [N,4]");
}
void check(lsd_status_t s) { if (s < 0) throw std::runtime_error(lsd_last_error()); }

farray to_numpy(const float* T) {
  farray out({4, 4});
  std::copy(T, T + 16, out.mutable_data());
  return out;
}

farray pointcloud_align(farray source_point, farray target_point, farray guess) {
  check_cloud(source_point, "source_point"); check_cloud(target_point, "target_point");
  if (guess.ndim() != 2 || guess.shape(0) != 4 || guess.shape(1) != 4) throw std::invalid_argument("guess must be [4,4]");
  float g[16];
  std::copy(guess.data(), guess.data() + 16, g);
  const double dist = std::sqrt((double)g[3] * g[3] + (double)g[7] * g[7] + (double)g[11] * g[11]);
  if (dist >= 50.0) g[3] = g[7] = g[11] = 0.f;                          // graph_utils.cpp:24-32
  lsd_reg_params_t p;
  lsd_reg_default_params(&p, LSD_REG_GICP);
  p.transformation_epsilon = 1e-2; p.max_iterations = 64; p.k_correspondences = 20;   // graph_utils.cpp:35-37
  p.max_corr_dist = 5.0;                                                // pcl::GeneralizedIterativeClosestPoint default gate
  Reg r(LSD_REG_GICP, p);
  float T[16];
  int conv = 0, it = 0;
  {
    py::gil_scoped_release nogil;
    check(lsd_reg_set_source(r.h, source_point.data(), (int)source_point.shape(0)));
    check(lsd_reg_set_target(r.h, target_point.data(), (int)target_point.shape(0)));
    check(lsd_reg_align(r.h, g, T, &conv, &it));
  }
  return to_numpy(T);
}

py::tuple registration_align(const std::string& method, farray source, farray target, farray guess, double max_corr,
                             long long max_process_time_us) {
  check_cloud(source, "source"); check_cloud(target, "target");
  if (guess.ndim() != 2 || guess.shape(0) != 4 || guess.shape(1) != 4) throw std::invalid_argument("guess must be [4,4]");
  int kind;
  if (method == "NDT_CUDA") kind = LSD_REG_NDT_P2D;                      // registrations.cpp:107-118
  else if (method == "FAST_GICP") kind = LSD_REG_GICP;                   // registrations.cpp:33-41
  else if (method == "FAST_VGICP" || method == "FAST_VGICP_CUDA") kind = LSD_REG_VGICP;  // registrations.cpp:44-66
  else throw std::invalid_argument("unknown registration method " + method);
  lsd_reg_params_t p;
  lsd_reg_default_params(&p, kind);
  if (max_corr > 0) p.max_corr_dist = max_corr;
  p.max_process_time_us = max_process_time_us;
  Reg r(kind, p);
  float T[16];
  int conv = 0, it = 0;
  double fit = 0;
  {
    py::gil_scoped_release nogil;
    check(lsd_reg_set_target(r.h, target.data(), (int)target.shape(0)));
    check(lsd_reg_set_source(r.h, source.data(), (int)source.shape(0)));
    check(lsd_reg_align(r.h, guess.data(), T, &conv, &it));
    check(lsd_reg_fitness(r.h, nullptr, 25.0, &fit));
  }
  return py::make_tuple(to_numpy(T), conv != 0, fit);
}

void dump_keyframe(const std::string& directory, uint64_t stamp, int id, farray points_input, farray pose_input) {
  if (points_input.ndim() != 2 || points_input.shape(1) != 4) throw std::invalid_argument("points_input must be float32 [N,4]");
  if (pose_input.ndim() != 2 || pose_input.shape(0) != 4 || pose_input.shape(1) != 4) throw std::invalid_argument("pose_input must be [4,4]");
  double pose[16];
  for (int i = 0; i < 16; i++) pose[i] = pose_input.data()[i];   // numpy_to_eigen: float -> double, py_utils.cpp
  const float* pts = points_input.data();
  const int n = (int)points_input.shape(0);
  lsd_status_t s;
  { py::gil_scoped_release release; s = lsd_keyframe_save(directory.c_str(), stamp, id, pts, n, pose); }
  check(s);
}

}  // namespace

PYBIND11_MODULE(slam_wrapper, m) {
  m.doc() = "mapping python interface (lsdreg hot-path subset)";
  m.def("pointcloud_align", &pointcloud_align, "pointcloud align", py::arg("source_point"), py::arg("target_point"), py::arg("guess"));
  m.def("registration_align", &registration_align, "select_registration_method(method) + align + fitness", py::arg("method"),
        py::arg("source"), py::arg("target"), py::arg("guess"), py::arg("max_corr") = 0.0, py::arg("max_process_time_us") = 0LL);
  m.def("dump_keyframe", &dump_keyframe, "dump keyframe", py::arg("directory"), py::arg("stamp"), py::arg("id"), py::arg("points_input"),
        py::arg("pose_input"));
  m.def("lsd_version", []() { return std::string(lsd_version()); });
}
