/* placeholder, filled in below */
#include "cc_oracle.h"
extern "C" int cco_mcmc(const cco_desc*, const double*, const cco_mcmc_opts*, double*, double*, double*, double*, double*) { return -1; }
