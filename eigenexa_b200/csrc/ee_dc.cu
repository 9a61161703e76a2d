// placeholder replaced below
#include "ee_common.cuh"
namespace ee {
int dc_dev(int, int, const double *, const double *, double *, double *, int) { set_error("dc not built"); return -1; }
}
