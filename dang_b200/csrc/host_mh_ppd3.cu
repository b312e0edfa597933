#define DG_PPD_BPL 3
#include "host_mh_ppd.inc"
