// ImageRegistration.h -- multi-resolution registration driver (reference src/ImageRegistration.h:11-58):
// owns the image / motion pyramids and one solver per level, runs the coarse-to-fine loop and leaves
// the per-level refine + iteration loop to the method family (OpticalFlow / Demons / Fluid).
#ifndef OF2D_HOST_IMAGE_REGISTRATION_H
#define OF2D_HOST_IMAGE_REGISTRATION_H

#include <vector>

#include <src/Image.h>
#include <src/Motion.h>
#include <src/SolverOptions.h>
#include <src/coord2d.h>
#include <src/regularization/IterativeSolver.h>

// what happened during the last estimate_motion() -- an extension for harnesses and tests
struct RegistrationTrace {
    struct Level {
        int scale = 0;
        int refine = 0;
        int iterations = 0;                 // get_update calls actually made
        std::vector<double> error;          // Logger error per iteration
        std::vector<int> regrid_iteration;  // Fluid only
        std::vector<double> regrid_minjac;
        std::vector<double> fluid_maxabs;   // Fluid only, per iteration
        std::vector<double> fluid_dt;
    };
    std::vector<Level> levels;
    long total_iterations() const {
        long n = 0;
        for (const Level& l : levels) n += l.iterations;
        return n;
    }
};

class ImageRegistration {
public:
    ImageRegistration(const dim dimin, const int nscales, const int* niter, const int nrefine, const Regularisation reg,
                      const of2d_real* regparams, const unsigned int nparams, const Verbose verbose);
    virtual ~ImageRegistration();   // virtual (the reference's is not: SURVEY Q16)

    void set_reference_image(const Image& im);
    void set_moving_image(const Image& im);
    Motion* get_estimated_motion() const;      // non-owning, valid until this object dies
    void copy_estimated_motion(Motion& mo) const;

    void estimate_motion();

    // extension: zero every level's motion and the solvers' carried state, so the next estimate_motion()
    // starts cold (the reference warm-starts from the previous result)
    void reset_state();

    const RegistrationTrace& get_trace() const { return trace; }

    // extension (pipelined sessions, of2d_sessions_register): the finest-level images as device arrays, to be filled on a copy
    // stream, and the coarser levels rebuilt from them (what set_reference_image / set_moving_image do after their copy)
    Image* reference_level0() { return Iref[0]; }
    Image* moving_level0() { return Imov[0]; }
    void rebuild_image_pyramids();

protected:
    void display_registration_parameters(const Regularisation reg, const of2d_real* regparams, const unsigned int nparams) const;

    virtual bool valid_regularisation_parameters(const Regularisation reg, const unsigned int nparams) const { return true; }
    virtual void set_solver(const Regularisation reg, const of2d_real* regparams, const unsigned int nparams) {}
    virtual void estimate_motion_at_current_resolution(Motion* motion, const Image* Iref, Image* Imov, IterativeSolver* solver, const int niter,
                                                       const dim dimin, const int sizein) {}

    // the refine + iteration loop shared by the three families (they differ in three places only)
    enum class LoopKind { OpticalFlow, Demons, Fluid };
    void run_level(LoopKind kind, Motion* motion, const Image* Iref, Image* Imov, IterativeSolver* solver, const int niter, const dim dimin);
    void release_solvers();

    // device-resident iteration engine (include/of2d_cuda.h), one per pyramid level, created on first use.
    // Used in the default (fast) mode; strict mode and parameter sets the engine refuses run the
    // per-iteration loop through the solvers' get_update().
    of2d_engine* engine_for_level(int level);
    bool run_level_on_engine(LoopKind kind, Motion* motion, const Image* Iref, Image* Imov, const int niter, const dim dimin);
    void release_engines();

    dim* dimin;
    int* sizein;
    int nscales;
    int* niter;
    int nrefine;

    IterativeSolver** solver;
    Image** Iref;
    Image** Imov;
    Motion** motion;

    Verbose verbose;
    RegistrationTrace trace;
    int current_scale = 0;

    Regularisation method;
    std::vector<of2d_real> method_params;
    std::vector<of2d_engine*> engines;        // [nscales + 1]
    std::vector<char> engine_refused;         // [nscales + 1]
};

#endif
