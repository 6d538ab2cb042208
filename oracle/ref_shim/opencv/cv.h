// Stand-in for <opencv/cv.h> (TEST INFRASTRUCTURE, see opencv2/core/core.hpp here).
// The real header pulls in OpenCV's C API (core/types_c.h), which includes the C headers <stdlib.h> and <math.h>.
// That matters to the reference: with libstdc++ >= 6 those wrappers put std::abs(double) into the global namespace, so
// the unqualified `abs(dist)` of matcher.cc:169 is a floating-point |dist| (GCC 5, the compiler of the reference's
// Ubuntu 16.04, has no such wrappers and would truncate dist to int there).  The stand-in keeps today's behaviour.
#pragma once
#include <stdlib.h>
#include <math.h>
#include <opencv2/core/core.hpp>
