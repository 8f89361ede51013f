/* oracle shim (test infrastructure): <cglm/vec3.h> forwards to the single-header restatement. */
#include "cglm.h"
