// Logger.h -- per-iteration convergence record (reference src/Logger.h:8-44): relative change
// mean||u_k - u_{k-1}|| / mean||u_{k-1}|| drives the early break of every driver loop.
#ifndef OF2D_HOST_LOGGER_H
#define OF2D_HOST_LOGGER_H

#include <vector>

#include <src/Motion.h>
#include <src/SolverOptions.h>
#include <src/coord2d.h>

class Logger {
public:
    Logger(const dim dimin, const unsigned int niter, const Verbose verbose);
    ~Logger();

    void update_error(const Motion* motion);
    of2d_real get_error_at_current_iteration() const;

    // extensions used by the harness / tests
    unsigned int iterations() const { return iter; }
    const std::vector<of2d_real>& errors() const { return error; }

private:
    void show_error_at_iteration(const unsigned int iter) const;
    void show_error_at_current_iteration() const;
    void show_all_error() const;

    dim dimin;
    unsigned int sizein;
    Motion prev;                       // u_{k-1}, device resident
    unsigned int niter;
    std::vector<of2d_real> error;      // niter + 1 slots
    unsigned int iter = 0;
    Verbose verbose;
};

#endif
