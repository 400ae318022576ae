// WrapperOpticalFlow2d.cpp -- MEX entry point `OpticalFlow2d(...)`, call-compatible with the
// reference's WrapperOpticalFlow2d.cpp:18-155.  One registration object lives between calls; the
// call shape (nlhs, nrhs) selects the action:
//   (0, 8)  create    prhs = {[dimx dimy], niter[nscales+1], nscales, reg, regparams[], nparams, nrefine, verbose}
//   (0, 2)  register  prhs = {Iref, Imov}            (double, dimx x dimy, column-major)
//   (1, 0)  motion    plhs[0] = double dimx x dimy x 2
//   (1, 1)  warp      plhs[0] = prhs[0] warped by the estimated motion
//   (0, 0)  destroy
// Everything numeric arrives as real doubles and is truncated to int / float as the reference does.
#include <memory>
#include <vector>

#include <mex.h>

#include <src/Image.h>
#include <src/ImageRegistrationDemons.h>
#include <src/ImageRegistrationFluid.h>
#include <src/ImageRegistrationOpticalFlow.h>
#include <src/Motion.h>
#include <src/SolverOptions.h>
#include <src/coord2d.h>

namespace {

struct Session {
    std::unique_ptr<ImageRegistration> registration;
    dim grid;
    mwSize image_dims[2];
    mwSize motion_dims[3];
};

std::unique_ptr<Session> g_session;

ImageRegistration* make_registration(dim grid, int nscales, const int* niter, int nrefine, Regularisation reg, const of2d_real* params,
                                     unsigned int nparams, Verbose verbose) {
    switch (reg) {
        case Regularisation::Diffusion:
        case Regularisation::Curvature:
        case Regularisation::Elastic:
            return new ImageRegistrationOpticalFlow(grid, nscales, niter, nrefine, reg, params, nparams, verbose);
        case Regularisation::ThirionsDemons:
        case Regularisation::DiffeomorphicDemons:
            return new ImageRegistrationDemons(grid, nscales, niter, nrefine, reg, params, nparams, verbose);
        case Regularisation::Fluid:
            return new ImageRegistrationFluid(grid, nscales, niter, nrefine, reg, params, nparams, verbose);
    }
    mexErrMsgTxt("Error: invalid regularisation given\n");
    return nullptr;
}

void create(const mxArray* prhs[]) {
    const double* dims = mxGetPr(prhs[0]);
    const int dimx = (int)dims[0], dimy = (int)dims[1];
    const int nscales = (int)mxGetPr(prhs[2])[0];
    const double* niter_d = mxGetPr(prhs[1]);
    std::vector<int> niter((size_t)nscales + 1);
    for (int s = 0; s <= nscales; s++) niter[(size_t)s] = (int)niter_d[s];
    const Regularisation reg = static_cast<Regularisation>((int)mxGetPr(prhs[3])[0]);
    const unsigned int nparams = (unsigned int)mxGetPr(prhs[5])[0];
    const double* params_d = mxGetPr(prhs[4]);
    std::vector<of2d_real> params(nparams ? nparams : 1);
    for (unsigned int p = 0; p < nparams; p++) params[p] = (of2d_real)params_d[p];
    const int nrefine = (int)mxGetPr(prhs[6])[0];
    const Verbose verbose = static_cast<Verbose>((int)mxGetPr(prhs[7])[0]);

    std::unique_ptr<Session> s(new Session());
    s->grid = dim((unsigned int)dimx, (unsigned int)dimy);
    s->registration.reset(make_registration(s->grid, nscales, niter.data(), nrefine, reg, params.data(), nparams, verbose));
    s->image_dims[0] = (mwSize)dimx; s->image_dims[1] = (mwSize)dimy;
    s->motion_dims[0] = (mwSize)dimx; s->motion_dims[1] = (mwSize)dimy; s->motion_dims[2] = 2;
    g_session = std::move(s);
}

}  // namespace

// used by the in-repo harness to read the trace of the live registration (NULL when none)
ImageRegistration* of2d_wrapper_registration() { return g_session ? g_session->registration.get() : nullptr; }

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const bool live = (bool)g_session;
    if (nlhs == 0 && nrhs == 8 && !live) {
        create(prhs);
    } else if (nlhs == 0 && nrhs == 2 && live) {
        Image reference(g_session->grid), moving(g_session->grid);
        reference.set_image(mxGetPr(prhs[0]));
        g_session->registration->set_reference_image(reference);
        moving.set_image(mxGetPr(prhs[1]));
        g_session->registration->set_moving_image(moving);
        g_session->registration->estimate_motion();
    } else if (nlhs == 1 && nrhs == 0 && live) {
        Motion result(g_session->grid);
        g_session->registration->copy_estimated_motion(result);
        plhs[0] = mxCreateNumericArray(3, g_session->motion_dims, mxDOUBLE_CLASS, mxREAL);
        result.copy_motion_to_input(mxGetPr(plhs[0]));
    } else if (nlhs == 1 && nrhs == 1 && live) {
        Image moving(g_session->grid);
        moving.set_image(mxGetPr(prhs[0]));
        moving.warp2d(*g_session->registration->get_estimated_motion());
        plhs[0] = mxCreateNumericArray(2, g_session->image_dims, mxDOUBLE_CLASS, mxREAL);
        moving.copy_image_to_input(mxGetPr(plhs[0]));
    } else if (nlhs == 0 && nrhs == 0 && live) {
        g_session.reset();
    } else {
        mexErrMsgTxt("Error: invalid number of input and output variables gives.\n");
    }
}
