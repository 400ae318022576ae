// DeviceRuntime.h -- the host classes' only door to the GPU: a process-wide of2d_ctx plus typed
// inline wrappers over the C ABI of include/of2d_cuda.h (float / double picked by overload).
// Nothing here includes CUDA headers; host code links libof2d_cuda.so and nothing else.
//
// Error mapping (same exception types the reference lets escape):
//   OF2D_ERR_INVALID -> std::invalid_argument     OF2D_ERR_DIVZERO -> std::runtime_error("Divide by zero exception")
//   OF2D_ERR_CUDA / UNSUPPORTED -> std::runtime_error(message)  (there is no CPU fallback)
#ifndef OF2D_HOST_DEVICE_RUNTIME_H
#define OF2D_HOST_DEVICE_RUNTIME_H

#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>

#include <of2d_cuda.h>
#include <src/coord2d.h>

namespace of2d {

of2d_ctx* context();            // created on first use on device $OF2D_DEVICE (else $LOCAL_RANK, else 0)
void release_context();         // tears the context down (tests)
// per-thread override of context(): the calling host thread works on `ctx` (and its device) until reset with nullptr;
// used by multi-GPU batches (one host thread + one context per device)
void set_thread_context(of2d_ctx* ctx);
void check(int status);         // throws per the table above
void poll_divzero();            // throws if a kernel raised the divide-by-zero flag since the last poll

// -------- device buffer with a lazily synchronised host mirror --------
// The reference API hands out raw mutable host pointers (get_image(), get_motion()); fields live in
// HBM, so the mirror is materialised on demand: host() downloads if stale and marks the device copy
// stale (the caller may write through the pointer); device accessors upload if needed.
class Buffer {
public:
    explicit Buffer(size_t bytes);
    Buffer(const Buffer& other);
    ~Buffer();
    Buffer& operator=(const Buffer&) = delete;

    size_t bytes() const { return bytes_; }
    const void* device_ro() const;
    void* device_rw();
    void* device_discard();       // contents will be fully overwritten on the device
    void* host() const;
    void zero();
    void copy_from(const Buffer& other);
    void swap(Buffer& other);

private:
    size_t bytes_;
    mutable void* dptr_;
    mutable void* hptr_;
    mutable bool device_valid_;
    mutable bool host_valid_;
};

}  // namespace of2d

// The wrappers below are plain overloads on the element type.
namespace of2d {
inline int image_from_double(size_t n, const double* in, float* out) { return of2d_image_from_double_f32(context(), n, in, out); }
inline int image_from_double(size_t n, const double* in, double* out) { return of2d_image_from_double_f64(context(), n, in, out); }
inline int image_to_double(size_t n, const float* in, double* out) { return of2d_image_to_double_f32(context(), n, in, out); }
inline int image_to_double(size_t n, const double* in, double* out) { return of2d_image_to_double_f64(context(), n, in, out); }
inline int motion_to_planar(size_t n, const float* u, double* out) { return of2d_motion_to_planar_double_f32(context(), n, u, out); }
inline int motion_to_planar(size_t n, const double* u, double* out) { return of2d_motion_to_planar_double_f64(context(), n, u, out); }
inline int warp2d(int nx, int ny, const float* s, const float* u, float* d) { return of2d_warp2d_f32(context(), nx, ny, 1, s, u, d); }
inline int warp2d(int nx, int ny, const double* s, const double* u, double* d) { return of2d_warp2d_f64(context(), nx, ny, 1, s, u, d); }
inline int compose(int nx, int ny, const float* u, const float* v, float* o) { return of2d_compose_f32(context(), nx, ny, 1, u, v, o); }
inline int compose(int nx, int ny, const double* u, const double* v, double* o) { return of2d_compose_f64(context(), nx, ny, 1, u, v, o); }
inline int convolute(int nc, int nx, int ny, const float* in, float* out, const double* k, int kw, int kh) {
    return nc == 2 ? of2d_convolute_motion_f32(context(), nx, ny, 1, in, out, k, kw, kh) : of2d_convolute_image_f32(context(), nx, ny, 1, in, out, k, kw, kh);
}
inline int convolute(int nc, int nx, int ny, const double* in, double* out, const double* k, int kw, int kh) {
    return nc == 2 ? of2d_convolute_motion_f64(context(), nx, ny, 1, in, out, k, kw, kh) : of2d_convolute_image_f64(context(), nx, ny, 1, in, out, k, kw, kh);
}
inline int derivatives(int nx, int ny, const float* r, const float* m, float* g, float* t) { return of2d_derivatives_f32(context(), nx, ny, 1, r, m, g, t); }
inline int derivatives(int nx, int ny, const double* r, const double* m, double* g, double* t) { return of2d_derivatives_f64(context(), nx, ny, 1, r, m, g, t); }
inline int jacobian(int nx, int ny, const float* u, float* j, float* mn) { return of2d_jacobian_f32(context(), nx, ny, u, j, mn); }
inline int jacobian(int nx, int ny, const double* u, double* j, double* mn) { return of2d_jacobian_f64(context(), nx, ny, u, j, mn); }
inline int axpy(size_t n, float a, const float* x, float* y) { return of2d_axpy_f32(context(), n, a, x, y); }
inline int axpy(size_t n, double a, const double* x, double* y) { return of2d_axpy_f64(context(), n, a, x, y); }
inline int scale(size_t n, float a, float* x) { return of2d_scale_f32(context(), n, a, x); }
inline int scale(size_t n, double a, double* x) { return of2d_scale_f64(context(), n, a, x); }
inline int scale_xy(size_t n, float ax, float ay, float* u) { return of2d_scale_xy_f32(context(), n, ax, ay, u); }
inline int scale_xy(size_t n, double ax, double ay, double* u) { return of2d_scale_xy_f64(context(), n, ax, ay, u); }
inline int motion_norm(size_t n, const float* u, float* h) { return of2d_motion_norm_f32(context(), n, u, h); }
inline int motion_norm(size_t n, const double* u, double* h) { return of2d_motion_norm_f64(context(), n, u, h); }
inline int motion_maxabs(size_t n, const float* u, float* h) { return of2d_motion_maxabs_f32(context(), n, u, h); }
inline int motion_maxabs(size_t n, const double* u, double* h) { return of2d_motion_maxabs_f64(context(), n, u, h); }
inline int image_stats(size_t n, const float* x, float* s, float* mx, float* mn) { return of2d_image_stats_f32(context(), n, x, s, mx, mn); }
inline int image_stats(size_t n, const double* x, double* s, double* mx, double* mn) { return of2d_image_stats_f64(context(), n, x, s, mx, mn); }
inline int image_normalize(size_t n, float lo, float hi, float* x) { return of2d_image_normalize_f32(context(), n, lo, hi, x); }
inline int image_normalize(size_t n, double lo, double hi, double* x) { return of2d_image_normalize_f64(context(), n, lo, hi, x); }
inline int motion_exp(int nx, int ny, float* u, float* tmp, int* ns) { return of2d_motion_exp_f32(context(), nx, ny, u, tmp, ns); }
inline int motion_exp(int nx, int ny, double* u, double* tmp, int* ns) { return of2d_motion_exp_f64(context(), nx, ny, u, tmp, ns); }
inline int downsample(int nc, int ix, int iy, const float* in, int ox, int oy, float* out) { return of2d_downsample_f32(context(), nc, ix, iy, in, ox, oy, out); }
inline int downsample(int nc, int ix, int iy, const double* in, int ox, int oy, double* out) { return of2d_downsample_f64(context(), nc, ix, iy, in, ox, oy, out); }
inline int upsample(int nc, int ix, int iy, const float* in, int ox, int oy, float* out) { return of2d_upsample_f32(context(), nc, ix, iy, in, ox, oy, out); }
inline int upsample(int nc, int ix, int iy, const double* in, int ox, int oy, double* out) { return of2d_upsample_f64(context(), nc, ix, iy, in, ox, oy, out); }
inline int boundary_conditions(int nx, int ny, int kind, float* u) { return of2d_boundary_conditions_f32(context(), nx, ny, kind, u); }
inline int boundary_conditions(int nx, int ny, int kind, double* u) { return of2d_boundary_conditions_f64(context(), nx, ny, kind, u); }
inline int logger_update(size_t n, const float* u, float* prev, float* h) { return of2d_logger_update_f32(context(), n, u, prev, h); }
inline int logger_update(size_t n, const double* u, double* prev, double* h) { return of2d_logger_update_f64(context(), n, u, prev, h); }
inline int lssd_force(int nx, int ny, const float* g, const float* t, const float* u, float* f) { return of2d_lssd_force_f32(context(), nx, ny, 1, g, t, u, f); }
inline int lssd_force(int nx, int ny, const double* g, const double* t, const double* u, double* f) { return of2d_lssd_force_f64(context(), nx, ny, 1, g, t, u, f); }
inline int diffusion_step(int nx, int ny, const float* u, float* un, const float* g, const float* t, float a) { return of2d_diffusion_step_f32(context(), nx, ny, 1, u, un, g, t, a, nullptr); }
inline int diffusion_step(int nx, int ny, const double* u, double* un, const double* g, const double* t, double a) { return of2d_diffusion_step_f64(context(), nx, ny, 1, u, un, g, t, a, nullptr); }
inline int elastic_step(int nx, int ny, float* u, const float* g, const float* t, float mu, float la, float om) { return of2d_elastic_step_f32(context(), nx, ny, 1, u, g, t, mu, la, om); }
inline int elastic_step(int nx, int ny, double* u, const double* g, const double* t, double mu, double la, double om) { return of2d_elastic_step_f64(context(), nx, ny, 1, u, g, t, mu, la, om); }
inline int fluid_step(int nx, int ny, float* u, float* v, float* r, const float* g, const float* t, float mu, float la, float om, float* hm, float* hd) { return of2d_fluid_step_f32(context(), nx, ny, u, v, r, g, t, mu, la, om, hm, hd); }
inline int fluid_step(int nx, int ny, double* u, double* v, double* r, const double* g, const double* t, double mu, double la, double om, double* hm, double* hd) { return of2d_fluid_step_f64(context(), nx, ny, u, v, r, g, t, mu, la, om, hm, hd); }
inline int demons_force(int nx, int ny, const float* r, const float* m, const float* u, float* c, float si, float sx) { return of2d_demons_force_f32(context(), nx, ny, 1, r, m, u, c, si, sx, nullptr); }
inline int demons_force(int nx, int ny, const double* r, const double* m, const double* u, double* c, double si, double sx) { return of2d_demons_force_f64(context(), nx, ny, 1, r, m, u, c, si, sx, nullptr); }
inline int demons_correspondence(size_t n, const float* g, const float* t, float* c, float si, float sx) { return of2d_demons_correspondence_f32(context(), n, g, t, c, si, sx); }
inline int demons_correspondence(size_t n, const double* g, const double* t, double* c, double si, double sx) { return of2d_demons_correspondence_f64(context(), n, g, t, c, si, sx); }
inline int curvature_step(of2d_curvature_plan* p, const float* u, float* un, const float* g, const float* t) { return of2d_curvature_step_f32(p, u, un, g, t); }
inline int curvature_step(of2d_curvature_plan* p, const double* u, double* un, const double* g, const double* t) { return of2d_curvature_step_f64(p, u, un, g, t); }
}  // namespace of2d

#endif
