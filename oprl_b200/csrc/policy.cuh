// SIMT kernels of the stochastic-policy algorithms (SAC / TQC): tanh-Gaussian head
// forward/backward, temperature step, TQC atom sort + quantile-Huber loss.
#pragma once
#include "kernels.cuh"

namespace oprl {}  // namespace oprl
