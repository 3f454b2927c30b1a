/* ORACLE TEST INFRASTRUCTURE. Empty stand-in: the reference's HDF5 output is `#if 0`
 * (src/output.c), only the include survives. */
