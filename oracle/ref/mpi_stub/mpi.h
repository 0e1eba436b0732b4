/* Single-rank stand-in for <mpi.h>, TEST INFRASTRUCTURE ONLY (oracle/): the reference's headers call
 * MPI_Comm_rank/size and a handful of collectives directly (CMakeLists.txt:87 requires MPI), and this
 * image has no MPI. With one rank every collective is a copy. Datatype handles are the element size in
 * bytes so the copies know how much to move. */
#ifndef HTB_ORACLE_MPI_STUB_H
#define HTB_ORACLE_MPI_STUB_H
#include <chrono>
#include <cstddef>
#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Fint;
#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_IN_PLACE ((void *)1)
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUCCESS 0
#define MPI_CHAR 1
#define MPI_UNSIGNED_CHAR 1
#define MPI_UNSIGNED_SHORT 2
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_UNSIGNED_LONG 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_C_COMPLEX 8
#define MPI_C_FLOAT_COMPLEX 8
#define MPI_C_DOUBLE_COMPLEX 16
#define MPI_DOUBLE_COMPLEX 16

inline int MPI_Init(int *, char ***) { return 0; }
inline int MPI_Finalize() { return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r) {
    *r = 0;
    return 0;
}
inline int MPI_Comm_size(MPI_Comm, int *s) {
    *s = 1;
    return 0;
}
inline int MPI_Barrier(MPI_Comm) { return 0; }
inline double MPI_Wtime() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline int MPI_Allgatherv(const void *s, int n, MPI_Datatype t, void *r, const int *, const int *d, MPI_Datatype, MPI_Comm) {
    if (s != MPI_IN_PLACE)
        std::memmove((char *)r + (size_t)d[0] * t, s, (size_t)n * t);
    return 0;
}
inline int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm) {
    if (s != MPI_IN_PLACE)
        std::memmove(r, s, (size_t)n * t);
    return 0;
}
inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
    if (s != MPI_IN_PLACE)
        std::memmove(r, s, (size_t)n * t);
    return 0;
}
inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
    if (s != MPI_IN_PLACE && s != r)
        std::memmove(r, s, (size_t)n * t);
    return 0;
}
inline int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype t, void *r, const int *, const int *rd, MPI_Datatype, MPI_Comm) {
    std::memmove((char *)r + (size_t)rd[0] * t, (const char *)s + (size_t)sd[0] * t, (size_t)sc[0] * t);
    return 0;
}
inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
#endif
