#define DG_PPD_BPL 2
#include "host_mh_ppd.inc"
