#define DG_PPD_BPL 5
#include "host_mh_ppd.inc"
