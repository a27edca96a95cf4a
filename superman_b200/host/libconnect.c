/* libConnect.so -- the shared object the reference's Python and MATLAB bindings load by name
 * (superPython.py:6-7: cdll.LoadLibrary('./libConnect.so'); libConnect.connect(); supermaTlab.m:
 * loadlibrary('libConnect', 'matlab_calculate_return.h')), with the four symbols of
 * interface_connector.c:61-231 under their reference names, forwarding to the GPU engine in
 * libsuperman_b200.so (host/sp_connector.c).  Built with -Wl,-Bsymbolic and meant to be dlopen'ed only:
 * its `connect` must never be linked into a program that also wants the socket call. */
#include "superman_b200.h"

double sp_read_calculate_return(char *filename, int algorithm, int nt, int x, int y, int z);
double sp_matlab_calculate_return_int(int *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz);
double sp_matlab_calculate_return_double(double *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz);

void connect(void) { sp_connect(); }                                       /* interface_connector.c:61 */

double read_calculate_return(char *filename, int algorithm, int nt, int x, int y, int z) {   /* :65 */
  return sp_read_calculate_return(filename, algorithm, nt, x, y, z);
}
double matlab_calculate_return_int(int *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz) {
  return sp_matlab_calculate_return_int(mat, algorithm, nt, x, y, z, nov, nnz);
}
double matlab_calculate_return_double(double *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz) {
  return sp_matlab_calculate_return_double(mat, algorithm, nt, x, y, z, nov, nnz);
}
