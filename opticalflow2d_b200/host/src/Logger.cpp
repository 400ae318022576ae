#include <src/Logger.h>

#include <mex.h>

Logger::Logger(const dim dimin_, const unsigned int niter_, const Verbose verbose_)
    : dimin(dimin_), sizein(dimin_.x * dimin_.y), prev(dimin_), niter(niter_), error(niter_ + 1, 0), verbose(verbose_) {}

Logger::~Logger() {}

// reference src/Logger.cpp:32-51.  One fused pass produces both sums and refreshes `prev`; the
// kernels' divide-by-zero flag rides along with the scalar read-back.
void Logger::update_error(const Motion* motion) {
    of2d_real e = 0;
    of2d::check(of2d::logger_update(sizein, motion->device(), prev.device_mut(), &e));
    if (iter < error.size()) error[iter] = e;
    else error.push_back(e);
    if (verbose == Verbose::On) show_error_at_current_iteration();
    iter++;
}

of2d_real Logger::get_error_at_current_iteration() const {
    if (iter == 0 || iter - 1 > niter)
        mexErrMsgTxt("Error: Logger::iter > Logger::niter, so current iteration cannot be shown,\n");
    return error[iter - 1];
}

void Logger::show_error_at_iteration(const unsigned int it) const {
    if (it <= niter) mexPrintf("Iteration: %d\tError:%.4f\n", it, error[it]);
    else mexErrMsgTxt("Error: Logger::iter > Logger::niter, so current iteration cannot be shown.\n");
}

void Logger::show_error_at_current_iteration() const { show_error_at_iteration(iter); }

void Logger::show_all_error() const {
    for (unsigned int it = 0; it <= iter; it++) show_error_at_iteration(it);
}
