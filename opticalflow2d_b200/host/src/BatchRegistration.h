// BatchRegistration.h -- extension: `batch` independent image pairs of one size registered together on
// one GPU (volume slices, cine frames; BASELINE.json configs[4]).  Every pair gets exactly what a fresh
// ImageRegistration{OpticalFlow,Demons,Fluid} object with nscales = 0 would compute for it (cold start:
// zero motion, zero fluid velocity, src/ImageRegistration.cpp:133-156 with one level), the pairs share
// the kernels of the device-resident iteration engine and carry their own control blocks, so pairs that
// converge early simply stop consuming work.  There is no exchange between pairs: multi-GPU runs shard
// the pair index across devices (one host thread + one context per GPU in this process: of2d_batch_create_multi,
// or one process per GPU).
//
// Cine chains (`frames` > 1): the batch is `frames` consecutive frame pairs of batch / frames independent
// sequences, frame-major (pair f * S + s is frame f of sequence s).  Frame f of a sequence starts from the motion
// -- and, for Fluid, the velocity -- frame f - 1 ended with, which is what the reference does when
// estimate_motion() is called again on the same object (src/ImageRegistration.cpp:135-139 does not reset
// motion[nscales]; OpticalFlowFluid keeps its velocity).  One wave = one frame.
#ifndef OF2D_HOST_BATCH_REGISTRATION_H
#define OF2D_HOST_BATCH_REGISTRATION_H

#include <vector>

#include <src/DeviceRuntime.h>
#include <src/SolverOptions.h>
#include <src/coord2d.h>

class BatchRegistration {
public:
    // regparams / nparams as for the single-pair classes (same validation); `wave`: pairs resident in the
    // engine at a time (0 = default 128; balanced over the waves, a partial last wave is padded internally);
    // `frames` > 1: cine chains (see above; batch must be a multiple of frames, wave = batch / frames)
    BatchRegistration(const dim dimin, const int batch, const int niter, const int nrefine, const Regularisation reg, const of2d_real* regparams,
                      const unsigned int nparams, const int wave = 0, const int frames = 1);
    ~BatchRegistration();
    BatchRegistration(const BatchRegistration&) = delete;
    BatchRegistration& operator=(const BatchRegistration&) = delete;

    // --- three-call protocol (images stay resident between estimates) ---
    // host doubles, batch images of dimx*dimy back to back (column-major each, as Image::set_image); returns when the
    // copies are complete (the caller may reuse its buffers)
    void set_images(const double* Iref, const double* Imov);
    void estimate_motion();
    // batch * 2 * dimx*dimy doubles: per pair the x plane then the y plane (Motion::copy_motion_to_input)
    void copy_estimated_motion(double* out) const;

    // --- streamed protocol: the three calls above in one, wave by wave, with the host -> device copy of wave k + 1 and
    // the device -> host copy of wave k - 1 running under the solve of wave k (two copy streams next to the compute
    // stream, double-buffered wave-sized device staging; the overlap needs pinned caller buffers) ---
    void register_pairs(const double* Iref, const double* Imov, double* planar_out);

    // per pair: iterations executed (summed over the refine passes), regrid events
    const std::vector<int>& iterations() const { return iters; }
    const std::vector<int>& regrids() const { return nregrid; }

    int size() const { return batch; }
    int wave_size() const { return wave; }
    int num_waves() const { return nwaves; }
    const of2d_real* device_motion() const { return (const of2d_real*)motion->device_ro(); }

private:
    void solve_wave(int w, const of2d_real* dr, const of2d_real* dm, of2d_real* mo, int pairs_in_wave);
    void pad_wave(of2d_ctx* ctx, of2d_real* img, int pairs_in_wave) const;

    dim grid;
    int batch, niter, nrefine, wave, nwaves, frames;
    size_t npix;
    of2d_engine* engine;
    of2d::Buffer *Iref, *Imov, *motion, *staging;
    // streamed protocol (created on first use)
    of2d_ctx *ctx_in, *ctx_out;
    of2d::Buffer *s_in[2], *s_ref[2], *s_mov[2], *s_mot[2], *s_out[2];
    std::vector<int> iters, nregrid;
};

#endif
