// 16-CTA-cluster decode kernel, configuration for 29..35 rows: three 3-warp attention groups per CTA, G <= 5.
#define H_NG 3
#define H_GW 3
#define H_GMAX 5
#define H_RING 3
#define H_KERNEL decode_mega16_kernel
#define H_CONFIGURE mega16_configure
#define H_LAUNCH mega16_launch
#include "mega16_impl.cuh"
