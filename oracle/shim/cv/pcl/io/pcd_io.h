// TEST INFRASTRUCTURE ONLY: see point_types.h.
#include "../point_types.h"
