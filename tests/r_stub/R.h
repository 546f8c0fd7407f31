/* TEST INFRASTRUCTURE - see Rinternals.h in this directory */
#ifndef ABN_R_STUB_R_H
#define ABN_R_STUB_R_H
#include <stdlib.h>
#include <string.h>
#endif
