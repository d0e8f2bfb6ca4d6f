"""TEST INFRASTRUCTURE: CPU restatement of the reference hot path and fixture generators (see oracle/c/oracle.h)."""
