#ifndef MD_ORACLE_H
#define MD_ORACLE_H
#endif
