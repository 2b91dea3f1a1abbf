/* fastpm_b200 -- the few MPI names the reference's Lua runtime helpers (src/lua-runtime-extended.c: os.get_nprocs) ask for,
 * answered by this build's communicator (one process per GPU; host/comm.c). */
#ifndef FASTPM_B200_LUA_MPI_H
#define FASTPM_B200_LUA_MPI_H
#include "fastpm_b200_api.h"
int fpm_comm_size(MPI_Comm comm);
static inline int MPI_Initialized(int *flag) { *flag = 1; return 0; }
#ifndef FASTPM_B200_HAVE_MPI_COMM_SIZE
static inline int fastpm_b200_lua_comm_size(MPI_Comm comm, int *np) { *np = fpm_comm_size(comm); return 0; }
#define MPI_Comm_size fastpm_b200_lua_comm_size
#endif
#endif
