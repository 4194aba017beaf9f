// gvalues.h -- numeric limits used by the matcher headers (reference: include/gmapping/utils/gvalues.h)
#ifndef GMB200_GVALUES_H
#define GMB200_GVALUES_H
#include <cfloat>
#include <climits>
#include <cmath>
#ifndef MAXDOUBLE
#define MAXDOUBLE DBL_MAX
#endif
#ifndef MINDOUBLE
#define MINDOUBLE DBL_MIN
#endif
#ifndef MAXFLOAT
#define MAXFLOAT FLT_MAX
#endif
#ifndef MINFLOAT
#define MINFLOAT FLT_MIN
#endif
#endif
