// temporary: entry points not implemented yet fail loudly (replaced as the kernels land)
#pragma once
#define RSD_NOT_YET(name) return rsd_fail(RSD_EINVAL, name ": not implemented in this build")
