// gplvm_dropin.h -- prefix header: builds the reference's UNMODIFIED gplvm.cpp front-end on CGplvmB200
// (`new CGplvm(&kern, &noise, latentDim, ...)` gplvm.cpp:522-536, `readGplvmFromFile` gplvm.cpp:643); see gp_dropin.h.
#ifndef GPLVM_DROPIN_H
#define GPLVM_DROPIN_H
#include "CGplvmB200.h"
#define CGplvm CGplvmB200
#define readGplvmFromFile readGplvmB200FromFile
#endif
