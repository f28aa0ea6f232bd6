// TEST INFRASTRUCTURE ONLY: see tbb.h in this directory.
#include "tbb.h"
