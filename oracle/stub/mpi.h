/* Test infrastructure (oracle/): a stand-in for <mpi.h> that lets the reference's OWN worker classes
 * (src/aslp-parallel/{bsp,bmuf,sod}-worker.cc over mpi-node.h) run N ranks as N threads of one process, so that their results can
 * pin the restated formulas (oracle/aslp_oracle.py) without an MPI installation.  Only what mpi-node.h:18-101 touches is declared;
 * the definitions live in oracle/ref_worker_driver.cc.  MPI_Allreduce adds the ranks' buffers in rank order, in the element type. */
#ifndef ASLP_ORACLE_STUB_MPI_H_
#define ASLP_ORACLE_STUB_MPI_H_
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_IN_PLACE ((void*)1)
#define MPI_SUM 0
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_FLOAT 3
#define MPI_DOUBLE 4
#define MPI_UNSIGNED 5
#define MPI_LONG_LONG_INT 6
int MPI_Init(int* argc, char*** argv);
int MPI_Finalize();
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
/* point to point, for the parameter-server modes ({easgd,asgd,masgd}-{server,worker}.cc): buffered sends into a per-rank
 * mailbox; a receive takes the first queued message that matches (source or MPI_ANY_SOURCE, tag or MPI_ANY_TAG), so messages
 * of one source are matched in the order they were sent, as MPI guarantees */
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, int source, int recvtag, MPI_Comm comm, MPI_Status* status);
#endif
