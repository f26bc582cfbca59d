/* forced include for oracle builds: declares the diagnostics sink the reference's PRNT macro is pointed at */
#ifndef ORACLE_PRNT_DECL_H
#define ORACLE_PRNT_DECL_H
int h_prnt(const char *fmt, ...);
#endif
