// 16-CTA-cluster decode kernel, configuration for <= 28 rows: two 4-warp attention groups per CTA, G <= 4.
#define H_NG 2
#define H_GW 4
#define H_GMAX 4
#define H_RING 6
#define H_KERNEL decode_mega16s_kernel
#define H_CONFIGURE mega16s_configure
#define H_LAUNCH mega16s_launch
#include "mega16_impl.cuh"
