/* tps_pgz.h -- parallel inflate of plain gzip input: declared in include/topsicle_host.h (part of the host ABI). */
#ifndef TPS_PGZ_H
#define TPS_PGZ_H
#include "../../include/topsicle_host.h"
#endif
