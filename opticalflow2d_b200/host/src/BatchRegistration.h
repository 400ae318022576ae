// BatchRegistration.h -- extension: `batch` independent image pairs of one size registered together on
// one GPU (volume slices, cine frames; BASELINE.json configs[4]).  Every pair gets exactly what a fresh
// ImageRegistration{OpticalFlow,Demons,Fluid} object with nscales = 0 would compute for it (cold start:
// zero motion, zero fluid velocity, src/ImageRegistration.cpp:133-156 with one level), the pairs share
// the kernels of the device-resident iteration engine and carry their own control blocks, so pairs that
// converge early simply stop consuming work.  There is no exchange between pairs: multi-GPU runs shard
// the pair index across processes (one per GPU).
#ifndef OF2D_HOST_BATCH_REGISTRATION_H
#define OF2D_HOST_BATCH_REGISTRATION_H

#include <vector>

#include <src/DeviceRuntime.h>
#include <src/SolverOptions.h>
#include <src/coord2d.h>

class BatchRegistration {
public:
    // regparams / nparams as for the single-pair classes (same validation); `wave`: pairs resident in the
    // engine at a time (0 = choose from the free device memory)
    BatchRegistration(const dim dimin, const int batch, const int niter, const int nrefine, const Regularisation reg, const of2d_real* regparams,
                      const unsigned int nparams, const int wave = 0);
    ~BatchRegistration();
    BatchRegistration(const BatchRegistration&) = delete;
    BatchRegistration& operator=(const BatchRegistration&) = delete;

    // host doubles, batch images of dimx*dimy back to back (column-major each, as Image::set_image)
    void set_images(const double* Iref, const double* Imov);
    void estimate_motion();
    // batch * 2 * dimx*dimy doubles: per pair the x plane then the y plane (Motion::copy_motion_to_input)
    void copy_estimated_motion(double* out) const;
    // per pair: iterations executed in the last refine, regrid events
    const std::vector<int>& iterations() const { return iters; }
    const std::vector<int>& regrids() const { return nregrid; }

    int size() const { return batch; }
    int wave_size() const { return wave; }
    const of2d_real* device_motion() const { return (const of2d_real*)motion->device_ro(); }

private:
    dim grid;
    int batch, niter, nrefine, wave;
    size_t npix;
    of2d_engine* engine;
    of2d::Buffer *Iref, *Imov, *motion, *staging;
    std::vector<int> iters, nregrid;
};

#endif
