/* Unit constants of the hot path.
 * QCxMS-side literals are byte-identical to the reference
 * (src/xtb_mctc_convert.f90:24-63, src/xtb_mctc_constants.f90:22-37, src/tblite.f90:43);
 * tblite-side conversion (radii tables in Angstrom -> bohr) uses CODATA-2018 like mctc-lib. */
#pragma once

/* QCxMS (old CODATA) */
#define QC_AUTOAA 0.52917726
#define QC_AATOAU (1.0 / QC_AUTOAA)
#define QC_AUTOEV 27.21138505
#define QC_EVTOAU (1.0 / QC_AUTOEV)
#define QC_AMUTOKG 1.660539040e-27
#define QC_METOKG 9.10938356e-31
#define QC_AMUTOAU (QC_AMUTOKG * (1.0 / QC_METOKG))
#define QC_AUTOAMU ((1.0 / QC_AMUTOKG) * QC_METOKG)
#define QC_FSTOAU 41.3413733365614
#define QC_KB 3.166808578545117e-06
#define QC_KTOAU 3.166808578545117e-06 /* src/tblite.f90:43 */
#define QC_MSTOAU (1.0 / 2.18769126364e+06)

/* tblite / mctc-lib side */
#define TB_AATOAU (1.0 / 0.529177210903)

/* method ids of the drop-in boundary (src/tblite.f90:34-40) */
#define QC_METHOD_GFN1 1
#define QC_METHOD_GFN2 2
#define QC_METHOD_IPEA1 11
/* stat codes (src/tblite.f90:25-27) */
#define QC_STAT_OK 0
#define QC_STAT_FATAL (-1)
#define QC_STAT_UNKNOWN_METHOD 5
