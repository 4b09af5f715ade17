// handle_registry.cpp — how many tracker handles are alive on a GPU, across PROCESSES.
//
// The per-frame path picks its kernel forms by that number: with one or two streams on a GPU most SMs idle during a frame and the
// "spread" forms trade them for latency; with more streams SM time is the budget (tracker_frame.cu: run_forward).  One process per
// stream (torchrun, one pipeline per camera process: /root/reference/src/main.rs runs one pipeline per process) must count like
// several streams in one process, so the count lives in a small POSIX shared-memory table keyed by the GPU's UUID: one slot per
// process {pid, handles}.  Slots of dead processes are reclaimed by whoever registers next (kill(pid, 0) == ESRCH).  If shared
// memory is unavailable the count degrades to the process-local one.
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <mutex>

#include "vt_internal.h"

namespace vt {

namespace {
constexpr int kSlots = 128;
struct Table {
    std::atomic<int32_t> pid[kSlots];
    std::atomic<int32_t> count[kSlots];
};
struct DeviceReg {
    Table* table = nullptr;   // shared mapping (null: process-local fallback)
    int slot = -1;            // this process's slot
    std::atomic<int> local{0};
    bool tried = false;
};
constexpr int kMaxDev = 64;
DeviceReg g_reg[kMaxDev];
std::mutex g_reg_mutex;

bool alive(int32_t pid) { return pid > 0 && (kill((pid_t)pid, 0) == 0 || errno != ESRCH); }

void attach(int device, DeviceReg& r) {
    r.tried = true;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    char name[96];
    const unsigned char* u = reinterpret_cast<const unsigned char*>(prop.uuid.bytes);
    snprintf(name, sizeof(name), "/vt_b200_%u_%02x%02x%02x%02x%02x%02x%02x%02x", (unsigned)getuid(), u[0], u[1], u[2], u[3], u[12], u[13], u[14], u[15]);
    const int fd = shm_open(name, O_RDWR | O_CREAT, 0600);
    if (fd < 0) return;
    if (ftruncate(fd, sizeof(Table)) != 0) {
        close(fd);
        return;
    }
    void* p = mmap(nullptr, sizeof(Table), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return;
    Table* tb = static_cast<Table*>(p);
    const int32_t me = (int32_t)getpid();
    for (int pass = 0; pass < 2 && r.slot < 0; ++pass)
        for (int i = 0; i < kSlots && r.slot < 0; ++i) {
            int32_t owner = tb->pid[i].load();
            if (owner == me) {  // a fork'd / re-attached process image: take it over
                r.slot = i;
            } else if (owner == 0 || (pass == 1 && !alive(owner))) {
                if (tb->pid[i].compare_exchange_strong(owner, me)) tb->count[i].store(0), r.slot = i;
            }
        }
    if (r.slot < 0) {
        munmap(p, sizeof(Table));
        return;
    }
    r.table = tb;
}
}  // namespace

// delta = +1 when a handle is created on `device`, -1 when it is destroyed
void registry_add(int device, int delta) {
    if (device < 0 || device >= kMaxDev) return;
    DeviceReg& r = g_reg[device];
    std::lock_guard<std::mutex> lock(g_reg_mutex);
    if (!r.tried) attach(device, r);
    const int now = r.local.fetch_add(delta) + delta;
    if (r.table) {
        r.table->count[r.slot].store(now);
        if (now == 0) r.table->pid[r.slot].store(0), r.slot = -1, munmap(r.table, sizeof(Table)), r.table = nullptr, r.tried = false;
    }
}

// handles alive on `device` in all processes of this user (read every frame: a sum over the table, no system call)
int registry_total(int device) {
    if (device < 0 || device >= kMaxDev) return 1;
    DeviceReg& r = g_reg[device];
    Table* tb = r.table;
    if (!tb) return r.local.load(std::memory_order_relaxed);
    int total = 0;
    for (int i = 0; i < kSlots; ++i)
        if (tb->pid[i].load(std::memory_order_relaxed) != 0) total += tb->count[i].load(std::memory_order_relaxed);
    return total > 0 ? total : r.local.load(std::memory_order_relaxed);
}

// drops the slots of processes that no longer exist (called when a handle is created: the frame path never makes a system call)
void registry_sweep(int device) {
    if (device < 0 || device >= kMaxDev) return;
    std::lock_guard<std::mutex> lock(g_reg_mutex);
    Table* tb = g_reg[device].table;
    if (!tb) return;
    for (int i = 0; i < kSlots; ++i) {
        int32_t owner = tb->pid[i].load();
        if (owner != 0 && !alive(owner) && tb->pid[i].compare_exchange_strong(owner, 0)) tb->count[i].store(0);
    }
}

}  // namespace vt
