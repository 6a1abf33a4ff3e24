// nccl_stub.cpp (cuemu) — TEST INFRASTRUCTURE: the eight NCCL entry points bendy2d_b200/csrc/solver.cu
// resolves with dlsym, implemented over named FIFOs between the rank processes of one machine, so that the
// multi-process strip path (StripSolver: one process per "GPU", halo exchange issued inside the captured
// graph) can run on the CPU emulation.  Send/Recv are handed to the emulated stream (cuemu_enqueue_host), so
// they are recorded during stream capture and replayed with the graph like the real ones.
#include <dlfcn.h>
#include <fcntl.h>
#include <poll.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
typedef enum { ncclSuccess = 0, ncclSystemError = 2, ncclInternalError = 3 } ncclResult_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
}

namespace {
struct Comm {
    int rank = 0, n = 0;
    std::string dir;
    std::vector<int> fd_out, fd_in;  // per peer
};
struct Op {
    bool send;
    char *buf;
    size_t bytes;
    int peer;
    Comm *comm;
};
thread_local std::vector<Op> g_group;
thread_local int g_depth = 0;

typedef int (*enqueue_fn)(void (*)(void *), void *);
enqueue_fn find_enqueue() {
    static enqueue_fn f = nullptr;
    if (f) return f;
    const char *lib = getenv("BENDY2D_B200_LIB");
    void *h = lib ? dlopen(lib, RTLD_NOW | RTLD_NOLOAD) : nullptr;
    if (h) f = (enqueue_fn)dlsym(h, "cuemu_enqueue_host");
    if (!f) f = (enqueue_fn)dlsym(RTLD_DEFAULT, "cuemu_enqueue_host");
    return f;
}

void run_ops(void *arg) {
    std::vector<Op> &ops = *static_cast<std::vector<Op> *>(arg);
    std::vector<size_t> done(ops.size(), 0);
    const auto t0 = std::chrono::steady_clock::now();
    while (true) {
        bool all = true, progress = false;
        std::vector<pollfd> pf;
        std::vector<int> busy;  // FIFOs that already have an unfinished earlier op: NCCL matches the sends and
                                // receives between two ranks in issue order, so a later op on the same FIFO waits
        for (size_t k = 0; k < ops.size(); k++) {
            Op &o = ops[k];
            if (done[k] == o.bytes) continue;
            all = false;
            const int fd = o.send ? o.comm->fd_out[o.peer] : o.comm->fd_in[o.peer];
            bool wait = false;
            for (int b : busy) wait = wait || b == fd;
            if (wait) continue;
            busy.push_back(fd);
            ssize_t r = o.send ? write(fd, o.buf + done[k], o.bytes - done[k]) : read(fd, o.buf + done[k], o.bytes - done[k]);
            if (r > 0) done[k] += (size_t)r, progress = true;
            pf.push_back(pollfd{fd, (short)(o.send ? POLLOUT : POLLIN), 0});
        }
        if (all) return;
        if (!progress) poll(pf.data(), pf.size(), 50);
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
            fprintf(stderr, "[cuemu nccl stub] rank %d: exchange timed out\n", ops[0].comm->rank);
            abort();
        }
    }
}

ncclResult_t flush_group() {
    if (g_group.empty()) return ncclSuccess;
    auto *ops = new std::vector<Op>(g_group);  // lives as long as a captured graph may replay it
    g_group.clear();
    enqueue_fn f = find_enqueue();
    if (!f) {
        fprintf(stderr, "[cuemu nccl stub] cuemu_enqueue_host not found (BENDY2D_B200_LIB must name the emulated library)\n");
        return ncclInternalError;
    }
    f(run_ops, ops);
    return ncclSuccess;
}
}  // namespace

extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
    memset(id, 0, sizeof *id);
    static int counter = 0;
    snprintf(id->internal, sizeof id->internal, "/tmp/cuemu_nccl_%d_%d_%ld", (int)getpid(), counter++,
             (long)std::chrono::steady_clock::now().time_since_epoch().count());
    return ncclSuccess;
}
ncclResult_t ncclCommInitRank(void **comm, int nranks, ncclUniqueId id, int rank) {
    Comm *c = new Comm();
    c->rank = rank, c->n = nranks, c->dir = id.internal;
    mkdir(c->dir.c_str(), 0700);
    c->fd_out.assign(nranks, -1), c->fd_in.assign(nranks, -1);
    for (int p = 0; p < nranks; p++) {
        if (p == rank) continue;
        const std::string out = c->dir + "/f_" + std::to_string(rank) + "_" + std::to_string(p);
        const std::string in = c->dir + "/f_" + std::to_string(p) + "_" + std::to_string(rank);
        mkfifo(out.c_str(), 0600);
        mkfifo(in.c_str(), 0600);
        c->fd_out[p] = open(out.c_str(), O_RDWR | O_NONBLOCK);  // O_RDWR on a FIFO never blocks in open
        c->fd_in[p] = open(in.c_str(), O_RDWR | O_NONBLOCK);
        if (c->fd_out[p] < 0 || c->fd_in[p] < 0) return ncclSystemError;
        fcntl(c->fd_out[p], F_SETPIPE_SZ, 1 << 20);
    }
    *comm = c;
    return ncclSuccess;
}
ncclResult_t ncclCommDestroy(void *comm) {
    Comm *c = static_cast<Comm *>(comm);
    for (int fd : c->fd_out)
        if (fd >= 0) close(fd);
    for (int fd : c->fd_in)
        if (fd >= 0) close(fd);
    return ncclSuccess;  // the Comm itself stays: a recorded graph may still point at it
}
static size_t dtype_size(int dt) { return (dt == 0 || dt == 1) ? 1 : (dt == 6 ? 2 : (dt == 4 || dt == 5 || dt == 8) ? 8 : 4); }
ncclResult_t ncclSend(const void *buf, size_t count, int dtype, int peer, void *comm, void *) {
    g_group.push_back(Op{true, (char *)buf, count * dtype_size(dtype), peer, static_cast<Comm *>(comm)});
    return g_depth ? ncclSuccess : flush_group();
}
ncclResult_t ncclRecv(void *buf, size_t count, int dtype, int peer, void *comm, void *) {
    g_group.push_back(Op{false, (char *)buf, count * dtype_size(dtype), peer, static_cast<Comm *>(comm)});
    return g_depth ? ncclSuccess : flush_group();
}
// sum of 64-bit integers over all ranks (the only reduction solver.cu asks for): everybody sends its vector
// to everybody, then adds what it received in rank order (integers: the order does not matter)
struct AllReduceOp {
    char *buf;
    size_t bytes;
    Comm *comm;
};
static void run_allreduce(void *arg) {
    AllReduceOp &a = *static_cast<AllReduceOp *>(arg);
    std::vector<std::vector<char>> in(a.comm->n, std::vector<char>(a.bytes));
    std::vector<char> mine(a.buf, a.buf + a.bytes);
    std::vector<Op> ops;
    for (int p = 0; p < a.comm->n; p++) {
        if (p == a.comm->rank) continue;
        ops.push_back(Op{true, mine.data(), a.bytes, p, a.comm});
        ops.push_back(Op{false, in[p].data(), a.bytes, p, a.comm});
    }
    run_ops(&ops);
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(a.buf);
    for (int p = 0; p < a.comm->n; p++) {
        if (p == a.comm->rank) continue;
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(in[p].data());
        for (size_t k = 0; k < a.bytes / 8; k++) dst[k] += src[k];
    }
}
ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, int dtype, int op, void *comm, void *) {
    if (send != recv || dtype_size(dtype) != 8 || op != 0) return ncclInternalError;  // in place, 64-bit, sum
    enqueue_fn f = find_enqueue();
    if (!f) return ncclInternalError;
    f(run_allreduce, new AllReduceOp{(char *)recv, count * 8, static_cast<Comm *>(comm)});
    return ncclSuccess;
}
ncclResult_t ncclGroupStart() {
    g_depth++;
    return ncclSuccess;
}
ncclResult_t ncclGroupEnd() {
    if (--g_depth > 0) return ncclSuccess;
    g_depth = 0;
    return flush_group();
}
const char *ncclGetErrorString(ncclResult_t) { return "cuemu nccl stub error"; }
}
