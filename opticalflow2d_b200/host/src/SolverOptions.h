// SolverOptions.h -- option enums of the registration API (values fixed by the MEX protocol,
// reference src/SolverOptions.h:4-8: the Octave caller passes them as plain numbers).
#ifndef OF2D_HOST_SOLVER_OPTIONS_H
#define OF2D_HOST_SOLVER_OPTIONS_H

enum Regularisation {
    Diffusion = 0,            // Horn-Schunck, Jacobi
    Curvature = 1,            // implicit curvature step, DCT
    Elastic = 2,              // Navier-Lame, one SOR sweep per iteration
    ThirionsDemons = 3,
    DiffeomorphicDemons = 4,
    Fluid = 5                 // viscous fluid with regridding
};

enum Verbose { Off = 0, On = 1 };

enum MotionAccumulation { Composition = 0, Addition = 1 };

#endif
