/* oracle/fake_r: see fake_r.h (test infrastructure only) */
#include "fake_r.h"
