#define DG_PPD_BPL 8
#include "host_mh_ppd.inc"
