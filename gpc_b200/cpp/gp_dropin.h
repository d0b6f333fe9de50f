// gp_dropin.h -- prefix header: builds the reference's UNMODIFIED front-end on the device-backed class.
//
//     g++ -std=gnu++98 -D_LINUX -I$GPC_REFERENCE -I$GPC_B200/include -I$GPC_B200/gpc_b200/cpp \
//         -include gp_dropin.h -c $GPC_REFERENCE/gp.cpp
//
// gp.cpp names the model class in three ways: `new CGp(&kern, &noise, &X, ...)` (gp.cpp:392), `CGp* pmodel` and
// `CGp::SCG`-style constants (gp.cpp:352-400), and `readGpFromFile(...)` (gp.cpp:487, 562, 619).  After CGp.h has been
// seen once (include guards make gp.cpp's own #include a no-op) the two names are redirected, so `gp learn`, `gp relearn`,
// `gp display` and `gp gnuplot` construct, optimise and query a CGpB200.
#ifndef GP_DROPIN_H
#define GP_DROPIN_H
#include "CGpB200.h"
#define CGp CGpB200
#define readGpFromFile readGpB200FromFile
#endif
