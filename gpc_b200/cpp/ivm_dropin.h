// ivm_dropin.h -- prefix header: builds the reference's UNMODIFIED ivm.cpp with its compound kernel on the device.
//
//     g++ -std=gnu++98 -D_LINUX -I$GPC_REFERENCE -I$GPC_B200/include -I$GPC_B200/gpc_b200/cpp \
//         -include ivm_dropin.h -c $GPC_REFERENCE/ivm.cpp
//
// ivm.cpp builds its kernel as `CCmpndKern kern(X)` (ivm.cpp:514) and hands it to CIvm as a CKern*; CIvm only calls the
// virtuals (CIvm.cpp:131 compute(K, X, X2) ...), so redirecting the one class name is enough.
#ifndef IVM_DROPIN_H
#define IVM_DROPIN_H
#include "CCmpndKernB200.h"
#define CCmpndKern CCmpndKernB200
#endif
