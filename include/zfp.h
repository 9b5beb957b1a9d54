/* zfp.h - compatibility shim: programs written against zfp's <zfp.h> compile against the
 * B200 backend's restatement of the array-path API (see zfp_b200.h). */
#ifndef ZFP_H
#define ZFP_H
#include "zfp_b200.h"
#endif
