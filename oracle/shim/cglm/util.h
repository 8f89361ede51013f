/* oracle shim (test infrastructure): <cglm/util.h> forwards to the single-header restatement. */
#include "cglm.h"
