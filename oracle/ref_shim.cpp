/*
 * ref_shim.cpp -- extern "C" entry points that drive the UNMODIFIED reference
 * sources (compiled where they lie under /root/reference by oracle/Makefile)
 * so that tests can call them through ctypes.  TEST INFRASTRUCTURE ONLY.
 *
 * Built twice: REAL=float against the reference as is, and REAL=double
 * against a build-time `sed s/float/double/` copy (the fp64 oracle of
 * SURVEY.md 8c).  Arrays crossing this boundary are REAL, column-major with x
 * fastest (idx = i + j*dimx, src/Field.tpp:13); motion fields are
 * array-of-structs {x,y} exactly as the reference stores them.
 */
#include <cstring>
#include <stdexcept>
#include <string>

#include <mex.h>

#include <src/Image.h>
#include <src/Kernel.h>
#include <src/Logger.h>
#include <src/Motion.h>
#include <src/SolverOptions.h>
#include <src/coord2d.h>
#include <src/regularization/Demons/DemonsDiffeomorphic.h>
#include <src/regularization/Demons/DemonsThirions.h>
#include <src/regularization/IterativeSolver.h>
#include <src/regularization/OpticalFlow/OpticalFlowCurvature.h>
#include <src/regularization/OpticalFlow/OpticalFlowDiffusion.h>
#include <src/regularization/OpticalFlow/OpticalFlowElastic.h>
#include <src/regularization/OpticalFlow/OpticalFlowFluid.h>

#ifndef OF2D_REF_REAL
#define OF2D_REF_REAL float
#endif
typedef OF2D_REF_REAL real;

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

namespace {

std::string g_last_error;

template <class F>
int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::invalid_argument &e) {
        g_last_error = e.what();
        return 2;
    } catch (const std::runtime_error &e) {
        g_last_error = e.what();
        return 3;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return 4;
    }
}

mxArray *make_mx(const double *vals, size_t n) {
    mwSize d[2] = {1, n};
    mxArray *a = mxCreateNumericArray(2, d, mxDOUBLE_CLASS, mxREAL);
    memcpy(a->data, vals, n * sizeof(double));
    return a;
}

void fill_image(Image &im, const real *src) { memcpy(im.get_image(), src, sizeof(real) * im.get_size()); }
void fill_motion(Motion &mo, const real *src) { memcpy(mo.get_motion(), src, 2 * sizeof(real) * mo.get_size()); }
void read_image(const Image &im, real *dst) { memcpy(dst, im.get_image(), sizeof(real) * im.get_size()); }
void read_motion(const Motion &mo, real *dst) { memcpy(dst, mo.get_motion(), 2 * sizeof(real) * mo.get_size()); }

/* the solvers keep gradI / It protected: a probe subclass reads them out */
struct DerivProbe : public IterativeSolver {
    explicit DerivProbe(dim d) : IterativeSolver(d) {}
    Motion *grad() { return gradI; }
    Image *dt() { return It; }
};

IterativeSolver *make_solver(int reg, dim d, const real *p, int np) {
    switch (reg) {
        case Diffusion: return new OpticalFlowDiffusion(d, p[0]);
        case Curvature: return np >= 2 ? new OpticalFlowCurvature(d, p[0], p[1]) : new OpticalFlowCurvature(d, p[0]);
        case Elastic: return np >= 3 ? new OpticalFlowElastic(d, p[0], p[1], p[2]) : new OpticalFlowElastic(d, p[0], p[1]);
        case ThirionsDemons:
            return new DemonsThirions(d, p[0], p[1], p[2], p[3], (unsigned int)p[4], static_cast<MotionAccumulation>((int)p[5]));
        case DiffeomorphicDemons: return new DemonsDiffeomorphic(d, p[0], p[1], p[2], p[3], (unsigned int)p[4]);
        case Fluid: return np >= 3 ? new OpticalFlowFluid(d, p[0], p[1], p[2]) : new OpticalFlowFluid(d, p[0], p[1]);
    }
    throw std::invalid_argument("unknown regularisation");
}

void delete_solver(int reg, IterativeSolver *s) {
    /* the reference's destructors are non-virtual (IterativeSolver.h:12): delete through the derived type */
    switch (reg) {
        case Diffusion: delete static_cast<OpticalFlowDiffusion *>(s); break;
        case Curvature: delete static_cast<OpticalFlowCurvature *>(s); break;
        case Elastic: delete static_cast<OpticalFlowElastic *>(s); break;
        case ThirionsDemons: delete static_cast<DemonsThirions *>(s); break;
        case DiffeomorphicDemons: delete static_cast<DemonsDiffeomorphic *>(s); break;
        case Fluid: delete static_cast<OpticalFlowFluid *>(s); break;
    }
}

}  // namespace

extern "C" {

int of2d_ref_sizeof_real(void) { return (int)sizeof(real); }
const char *of2d_ref_last_error(void) { return g_last_error.c_str(); }

/* ---- the 5-call MEX protocol of test_opticalflow2d.m:42-59, through the reference's own mexFunction ---- */
int of2d_ref_mex_register(int dimx, int dimy, int nscales, const double *niter, int nrefine, int reg,
                          const double *regparams, int nparams, int verbose, const double *Iref,
                          const double *Imov, double *motion_out, double *warped_out) {
    return guarded([&] {
        const size_t n = (size_t)dimx * dimy;
        double dims[2] = {(double)dimx, (double)dimy};
        double dnscales = nscales, dreg = reg, dnparams = nparams, dnrefine = nrefine, dverbose = verbose;
        mxArray *init[8] = {make_mx(dims, 2), make_mx(niter, nscales + 1), make_mx(&dnscales, 1),
                            make_mx(&dreg, 1), make_mx(regparams, nparams), make_mx(&dnparams, 1),
                            make_mx(&dnrefine, 1), make_mx(&dverbose, 1)};
        mxArray *imgs[2] = {make_mx(Iref, n), make_mx(Imov, n)};
        mxArray *out[1] = {NULL};
        auto cleanup = [&] {
            for (auto *a : init) mxDestroyArray(a);
            for (auto *a : imgs) mxDestroyArray(a);
        };
        try {
            mexFunction(0, NULL, 8, (const mxArray **)init);
            try {
                mexFunction(0, NULL, 2, (const mxArray **)imgs);
                mexFunction(1, out, 0, NULL);
                if (motion_out) memcpy(motion_out, out[0]->data, 2 * n * sizeof(double));
                mxDestroyArray(out[0]);
                out[0] = NULL;
                mexFunction(1, out, 1, (const mxArray **)&imgs[1]);
                if (warped_out) memcpy(warped_out, out[0]->data, n * sizeof(double));
                mxDestroyArray(out[0]);
            } catch (...) {
                mexFunction(0, NULL, 0, NULL); /* always release the singleton */
                throw;
            }
            mexFunction(0, NULL, 0, NULL);
        } catch (...) {
            cleanup();
            throw;
        }
        cleanup();
    });
}

/* raw mexFunction call shape probe: returns the status of an arbitrary (nlhs, nrhs) call with dummy args */
int of2d_ref_mex_badcall(int nlhs, int nrhs) {
    return guarded([&] {
        double z = 0;
        mxArray *args[8];
        for (int i = 0; i < 8; i++) args[i] = make_mx(&z, 1);
        mxArray *out[1] = {NULL};
        try {
            mexFunction(nlhs, out, nrhs, (const mxArray **)args);
        } catch (...) {
            for (int i = 0; i < 8; i++) mxDestroyArray(args[i]);
            throw;
        }
        for (int i = 0; i < 8; i++) mxDestroyArray(args[i]);
    });
}

/* ---- primitives ---- */
int of2d_ref_set_image(int dimx, int dimy, const double *in, real *out) {
    return guarded([&] { Image im(dim(dimx, dimy)); im.set_image(in); read_image(im, out); });
}
int of2d_ref_copy_motion_to_input(int dimx, int dimy, const real *u, double *out) {
    return guarded([&] { Motion mo(dim(dimx, dimy)); fill_motion(mo, u); mo.copy_motion_to_input(out); });
}
int of2d_ref_warp2d(int dimx, int dimy, real *img, const real *u) {
    return guarded([&] {
        dim d(dimx, dimy);
        Image im(d); Motion mo(d);
        fill_image(im, img); fill_motion(mo, u);
        im.warp2d(mo);
        read_image(im, img);
    });
}
int of2d_ref_accumulate(int dimx, int dimy, real *u, const real *v) {
    return guarded([&] {
        dim d(dimx, dimy);
        Motion a(d), b(d);
        fill_motion(a, u); fill_motion(b, v);
        a.accumulate(b);
        read_motion(a, u);
    });
}
int of2d_ref_gaussian_kernel(int w, real sigma, double *out) {
    return guarded([&] {
        Kernel k((unsigned int)w);
        k.set_gaussian(sigma);
        memcpy(out, k.get_kernel(), sizeof(double) * k.get_size());
    });
}
int of2d_ref_convolute_motion(int dimx, int dimy, real *u, int w, real sigma) {
    return guarded([&] {
        Motion a(dim(dimx, dimy));
        fill_motion(a, u);
        Kernel k((unsigned int)w);
        k.set_gaussian(sigma);
        a.convolute(k);
        read_motion(a, u);
    });
}
int of2d_ref_exp(int dimx, int dimy, real *u) {
    return guarded([&] { Motion a(dim(dimx, dimy)); fill_motion(a, u); a.exp(); read_motion(a, u); });
}
int of2d_ref_norm_maxabs(int dimx, int dimy, const real *u, real *norm, real *maxabs) {
    return guarded([&] { Motion a(dim(dimx, dimy)); fill_motion(a, u); *norm = a.norm(); *maxabs = a.maxabs(); });
}
int of2d_ref_jacobian(int dimx, int dimy, const real *u, real *jac, real *minjac) {
    return guarded([&] {
        dim d(dimx, dimy);
        Motion a(d); Image j(d);
        fill_motion(a, u);
        j.jacobian(a);
        read_image(j, jac);
        *minjac = j.min();
    });
}
int of2d_ref_derivatives(int dimx, int dimy, const real *Iref, const real *Imov, real *grad, real *It) {
    return guarded([&] {
        dim d(dimx, dimy);
        Image r(d), m(d);
        fill_image(r, Iref); fill_image(m, Imov);
        DerivProbe p(d);
        p.set_derivatives(&r, &m);
        read_motion(*p.grad(), grad);
        read_image(*p.dt(), It);
    });
}
int of2d_ref_image_resample(int inx, int iny, const real *in, int outx, int outy, real *out, int up) {
    return guarded([&] {
        Image a(dim(inx, iny)), b(dim(outx, outy));
        fill_image(a, in); fill_image(b, out);
        if (up) b.upSample(a); else b.downSample(a);
        read_image(b, out);
    });
}
int of2d_ref_motion_resample(int inx, int iny, const real *in, int outx, int outy, real *out, int up) {
    return guarded([&] {
        Motion a(dim(inx, iny)), b(dim(outx, outy));
        fill_motion(a, in); fill_motion(b, out);
        if (up) b.upSample(a); else b.downSample(a);
        read_motion(b, out);
    });
}
/* ---- the rest of the public Image / Motion / Kernel surface ---- */
int of2d_ref_image_stats(int dimx, int dimy, const real *img, real *sum, real *mx, real *mn) {
    return guarded([&] { Image a(dim(dimx, dimy)); fill_image(a, img); *sum = a.sum(); *mx = a.max(); *mn = a.min(); });
}
int of2d_ref_image_normalize(int dimx, int dimy, real *img) {
    return guarded([&] { Image a(dim(dimx, dimy)); fill_image(a, img); a.normalize(); read_image(a, img); });
}
int of2d_ref_boundary_conditions(int dimx, int dimy, int kind, real *u) {
    return guarded([&] {
        Motion a(dim(dimx, dimy));
        fill_motion(a, u);
        if (kind) a.Dirichlet_boundaryconditions(); else a.Neumann_boundaryconditions();
        read_motion(a, u);
    });
}
int of2d_ref_average_kernel(int w, double *out) {
    return guarded([&] { Kernel k((unsigned int)w); k.set_average(); memcpy(out, k.get_kernel(), sizeof(double) * k.get_size()); });
}
/* Image::convolute with a Gaussian (sigma > 0) or the average kernel (sigma <= 0).  Field<float>::convolute leaves its
   accumulator `val` uninitialised (src/Field.tpp:240): what this returns is whatever the compiled code does with it. */
int of2d_ref_convolute_image(int dimx, int dimy, real *img, int w, real sigma) {
    return guarded([&] {
        Image a(dim(dimx, dimy));
        fill_image(a, img);
        Kernel k((unsigned int)w);
        if (sigma > 0) k.set_gaussian(sigma); else k.set_average();
        a.convolute(k);
        read_image(a, img);
    });
}
/* Logger::update_error over a sequence of nseq fields (each 2*N reals); err has nseq entries */
int of2d_ref_logger(int dimx, int dimy, const real *useq, int nseq, real *err) {
    return guarded([&] {
        dim d(dimx, dimy);
        Logger log(d, nseq, Verbose::Off);
        Motion a(d);
        for (int s = 0; s < nseq; s++) {
            fill_motion(a, useq + (size_t)s * 2 * dimx * dimy);
            log.update_error(&a);
            err[s] = log.get_error_at_current_iteration();
        }
    });
}

/* nsteps calls of solver->get_update on state u (in/out), after set_derivatives where the driver does it.
   For Demons (reg 3,4) Imov is the moving image handed to get_update (ImageRegistrationDemons.cpp:111). */
int of2d_ref_solver_steps(int reg, const real *params, int nparams, int dimx, int dimy, const real *Iref,
                          const real *Imov, real *u, int nsteps) {
    return guarded([&] {
        dim d(dimx, dimy);
        Image r(d), m(d);
        Motion mo(d);
        fill_image(r, Iref); fill_image(m, Imov); fill_motion(mo, u);
        IterativeSolver *s = make_solver(reg, d, params, nparams);
        try {
            const bool demons = (reg == ThirionsDemons || reg == DiffeomorphicDemons);
            if (!demons) s->set_derivatives(&r, &m);
            for (int k = 0; k < nsteps; k++) s->get_update(&mo, &r, &m);
        } catch (...) {
            delete_solver(reg, s);
            throw;
        }
        delete_solver(reg, s);
        read_motion(mo, u);
    });
}

}  // extern "C"
