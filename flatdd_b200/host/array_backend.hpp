// array_backend.hpp — what the simulator driver talks to once the state has left the DD.
//
// Three implementations:
//   * GpuArrayBackend   — the product: forwards to the C-ABI in include/flatdd_b200.h
//                         (sm_100a kernels).  Throws std::runtime_error on any failure; there
//                         is no CPU fallback.
//   * TraceRecorder     — writes every flat table that crosses the boundary to a binary trace
//                         (used to ship a circuit's array phase to a box that has no
//                         reference tree, and to generate golden fixtures).
//   * TeeBackend        — fan-out to several backends.
// (The CPU backend that drives the *reference's own* DDArrMultiplyIP lives in
//  oracle/ref_dump.cpp: it is test infrastructure, not product.)
#pragma once

#include "flatdd_b200.h"
#include "flatten.hpp"

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace fddb200 {

class ArrayBackend {
public:
    virtual ~ArrayBackend() = default;
    // DD -> array (reference: getVectorFromDDSwitch1 / getVectorFromDD)
    virtual void convert(const FlatVecDD& dd) = 0;
    // one DMAVM (reference: DDArrMultiplyIP / DDArrMultiplyOP); `nOriginalGates` = how many
    // circuit operations were fused into this matrix (bookkeeping only)
    virtual void apply(const FlatMatDD& gate, int nOriginalGates) = 0;
    // a stretch of the schedule in one boundary call (reference: the executor loop, src/SwitchSimulator.cpp:386-412); the GPU
    // backend keeps the state tile-resident across consecutive dense blocks.  Default: one apply per gate.
    virtual void applyMany(const std::vector<FlatMatDD>& gates, const std::vector<int>& nOriginalGates) {
        for (std::size_t i = 0; i < gates.size(); ++i) apply(gates[i], nOriginalGates[i]);
    }
    // copy the current state into SoA arrays of 2^n_local doubles (reference: getVector)
    virtual void getState(double* real, double* imag) = 0;
    virtual void synchronize() {}
    // sharded states only: SWAP of a global and a local physical index bit (half-shard exchange),
    // and the undo of all such swaps so that shard r again holds the amplitudes with top bits r
    virtual void exchange(int /*globalPhysicalBit*/, int /*localPhysicalBit*/) {
        throw std::runtime_error("this backend holds an unsharded state");
    }
    // a stretch of the schedule and the exchange that follows it in one boundary call: the GPU backend lets the stretch's last
    // pass store the traded half straight into the partner shard (fdd_apply_many_exchange).  Default: the two calls.
    virtual void applyManyThenExchange(const std::vector<FlatMatDD>& gates, const std::vector<int>& nOriginalGates, int globalPhysicalBit,
                                       int localPhysicalBit) {
        applyMany(gates, nOriginalGates);
        exchange(globalPhysicalBit, localPhysicalBit);
    }
    // device milliseconds of the last convert / apply when per-launch timing is on (0 otherwise)
    virtual double lastKernelMs() { return 0.0; }
    // the logical qubits at two physical bits trade names (absorbed SWAP gate); no data moves
    virtual void relabel(int /*physicalBitA*/, int /*physicalBitB*/) {}
    virtual void canonicalize() {}
};

inline void fddCheck(int rc, const char* what) {
    if (rc != FDD_OK) {
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + fdd_last_error());
    }
}

class GpuArrayBackend final : public ArrayBackend {
public:
    GpuArrayBackend(int nQubits, int device = 0) { fddCheck(fdd_create(nQubits, device, &ctx_), "fdd_create"); }
    // one shard of a state distributed over `worldSize` processes; `ncclUniqueId128` comes from
    // fdd_comm_unique_id on rank 0 and reaches the other ranks through the launcher
    GpuArrayBackend(int nQubits, int device, int rank, int worldSize, const void* ncclUniqueId128, int exchangeMethod = 0)
        : exchangeMethod_(exchangeMethod) {
        fddCheck(fdd_create_sharded(nQubits, device, rank, worldSize, &ctx_), "fdd_create_sharded");
        if (worldSize > 1) {
            fddCheck(fdd_comm_init(ctx_, ncclUniqueId128), "fdd_comm_init");
        }
    }
    GpuArrayBackend(const GpuArrayBackend&) = delete;
    GpuArrayBackend& operator=(const GpuArrayBackend&) = delete;
    ~GpuArrayBackend() override {
        if (ctx_ != nullptr) {
            fdd_destroy(ctx_);
        }
    }
    void convert(const FlatVecDD& dd) override {
        const fdd_vecdd v = view(dd);
        fddCheck(fdd_convert(ctx_, &v), "fdd_convert");
    }
    void apply(const FlatMatDD& gate, int /*nOriginalGates*/) override {
        const fdd_matdd m = view(gate);
        fddCheck(fdd_apply(ctx_, &m), "fdd_apply");
    }
    void applyMany(const std::vector<FlatMatDD>& gates, const std::vector<int>& /*nOriginalGates*/) override {
        std::vector<fdd_matdd> views;
        views.reserve(gates.size());
        for (const auto& g : gates) views.push_back(view(g));
        fddCheck(fdd_apply_many(ctx_, views.data(), static_cast<int>(views.size())), "fdd_apply_many");
    }
    void getState(double* real, double* imag) override { fddCheck(fdd_get_state(ctx_, real, imag), "fdd_get_state"); }
    void synchronize() override { fddCheck(fdd_synchronize(ctx_), "fdd_synchronize"); }
    void exchange(int globalPhysicalBit, int localPhysicalBit) override {
        fddCheck(fdd_exchange_qubits(ctx_, globalPhysicalBit, localPhysicalBit, exchangeMethod_), "fdd_exchange_qubits");
    }
    void applyManyThenExchange(const std::vector<FlatMatDD>& gates, const std::vector<int>& nOriginalGates, int globalPhysicalBit,
                               int localPhysicalBit) override {
        if (exchangeMethod_ != 0) { // the fused exchange is the peer-memory path
            ArrayBackend::applyManyThenExchange(gates, nOriginalGates, globalPhysicalBit, localPhysicalBit);
            return;
        }
        std::vector<fdd_matdd> views;
        views.reserve(gates.size());
        for (const auto& g : gates) views.push_back(view(g));
        fddCheck(fdd_apply_many_exchange(ctx_, views.data(), static_cast<int>(views.size()), globalPhysicalBit, localPhysicalBit), "fdd_apply_many_exchange");
    }
    void relabel(int a, int b) override { fddCheck(fdd_relabel_qubits(ctx_, a, b), "fdd_relabel_qubits"); }
    void canonicalize() override { fddCheck(fdd_canonicalize(ctx_), "fdd_canonicalize"); }
    void setTiming(bool on) { fddCheck(fdd_set_timing(ctx_, on ? 1 : 0), "fdd_set_timing"); }
    double lastKernelMs() override {
        float ms = 0.0F;
        fddCheck(fdd_last_kernel_ms(ctx_, &ms), "fdd_last_kernel_ms");
        return static_cast<double>(ms);
    }
    [[nodiscard]] fdd_ctx* ctx() const { return ctx_; }

private:
    fdd_ctx* ctx_ = nullptr;
    int exchangeMethod_ = 0;
};

// Binary trace, little endian:
//   char[8] "FDDTRC01"; int32 n_qubits; int32 n_records (patched on close);
//   records: int32 kind (1 = vector DD to convert, 2 = matrix DD to apply), int32 n_nodes,
//            int32 root, int32 n_original_gates, double root_weight[2],
//            int32 level[n_nodes], int32 child[R*n_nodes], double weight[2*R*n_nodes]   (R = 2 or 4)
//            kind 3 = exchange: the n_nodes field holds the global physical bit, the root field the
//            local physical bit, no tables; kind 4 = relabel (two physical bits, same fields);
//            kind 5 = meta: the n_nodes field holds the world size
class TraceRecorder final : public ArrayBackend {
public:
    TraceRecorder(const std::string& path, int nQubits) : file_(std::fopen(path.c_str(), "wb")) {
        if (file_ == nullptr) {
            throw std::runtime_error("TraceRecorder: cannot open " + path);
        }
        const char magic[8] = {'F', 'D', 'D', 'T', 'R', 'C', '0', '1'};
        put(magic, sizeof magic);
        const int32_t n = nQubits;
        put(&n, sizeof n);
        put(&records_, sizeof records_);
    }
    TraceRecorder(const TraceRecorder&) = delete;
    TraceRecorder& operator=(const TraceRecorder&) = delete;
    ~TraceRecorder() override { close(); }
    void convert(const FlatVecDD& dd) override { record<2>(1, dd, 0); }
    void apply(const FlatMatDD& gate, int nOriginalGates) override { record<4>(2, gate, nOriginalGates); }
    void getState(double*, double*) override { throw std::runtime_error("TraceRecorder holds no state"); }
    void exchange(int globalPhysicalBit, int localPhysicalBit) override { marker(3, globalPhysicalBit, localPhysicalBit); }
    void relabel(int a, int b) override { marker(4, a, b); }
    void setWorldSize(int worldSize) { marker(5, worldSize, 0); }
    void close() {
        if (file_ != nullptr) {
            std::fseek(file_, 12, SEEK_SET);
            put(&records_, sizeof records_);
            std::fclose(file_);
            file_ = nullptr;
        }
    }
    [[nodiscard]] int32_t records() const { return records_; }

private:
    void put(const void* p, std::size_t bytes) {
        if (bytes != 0 && std::fwrite(p, 1, bytes, file_) != bytes) {
            throw std::runtime_error("TraceRecorder: short write");
        }
    }
    void marker(int32_t kind, int32_t a, int32_t b) {
        const int32_t head[4] = {kind, a, b, 0};
        const double zero[2] = {0.0, 0.0};
        put(head, sizeof head);
        put(zero, sizeof zero);
        ++records_;
    }
    template <int R> void record(int32_t kind, const FlatDD<R>& dd, int32_t nOriginal) {
        const int32_t head[4] = {kind, dd.nNodes(), dd.root, nOriginal};
        put(head, sizeof head);
        put(dd.root_weight, sizeof dd.root_weight);
        put(dd.level.data(), dd.level.size() * sizeof(int32_t));
        put(dd.child.data(), dd.child.size() * sizeof(int32_t));
        put(dd.weight.data(), dd.weight.size() * sizeof(double));
        ++records_;
    }
    std::FILE* file_;
    int32_t records_ = 0;
};

// Queues consecutive apply() calls and hands them to the wrapped backend in one applyMany() (so the GPU library can keep the state
// tile-resident across consecutive dense blocks); anything else flushes the queue first.
class BatchingBackend final : public ArrayBackend {
public:
    explicit BatchingBackend(ArrayBackend* inner, std::size_t maxQueued = 64) : inner_(inner), maxQueued_(maxQueued) {}
    ~BatchingBackend() override {
        try {
            flush();
        } catch (...) {
        }
    }
    void convert(const FlatVecDD& dd) override {
        flush();
        inner_->convert(dd);
    }
    void apply(const FlatMatDD& gate, int nOriginalGates) override {
        gates_.push_back(gate);
        originals_.push_back(nOriginalGates);
        if (gates_.size() >= maxQueued_) flush();
    }
    void applyMany(const std::vector<FlatMatDD>& gates, const std::vector<int>& nOriginalGates) override {
        flush();
        inner_->applyMany(gates, nOriginalGates);
    }
    void getState(double* real, double* imag) override {
        flush();
        inner_->getState(real, imag);
    }
    void synchronize() override {
        flush();
        inner_->synchronize();
    }
    void exchange(int globalPhysicalBit, int localPhysicalBit) override {
        if (gates_.empty()) {
            inner_->exchange(globalPhysicalBit, localPhysicalBit);
            return;
        }
        inner_->applyManyThenExchange(gates_, originals_, globalPhysicalBit, localPhysicalBit); // the queued stretch takes the exchange along
        gates_.clear();
        originals_.clear();
    }
    void applyManyThenExchange(const std::vector<FlatMatDD>& gates, const std::vector<int>& nOriginalGates, int globalPhysicalBit,
                               int localPhysicalBit) override {
        flush();
        inner_->applyManyThenExchange(gates, nOriginalGates, globalPhysicalBit, localPhysicalBit);
    }
    void relabel(int a, int b) override {
        flush();
        inner_->relabel(a, b);
    }
    void canonicalize() override {
        flush();
        inner_->canonicalize();
    }
    double lastKernelMs() override { return inner_->lastKernelMs(); }
    void flush() {
        if (gates_.empty()) return;
        inner_->applyMany(gates_, originals_);
        gates_.clear();
        originals_.clear();
    }

private:
    ArrayBackend* inner_;
    std::size_t maxQueued_;
    std::vector<FlatMatDD> gates_;
    std::vector<int> originals_;
};

class TeeBackend final : public ArrayBackend {
public:
    explicit TeeBackend(std::vector<ArrayBackend*> sinks) : sinks_(std::move(sinks)) {}
    void convert(const FlatVecDD& dd) override {
        for (auto* s : sinks_) {
            s->convert(dd);
        }
    }
    void apply(const FlatMatDD& gate, int nOriginalGates) override {
        for (auto* s : sinks_) {
            s->apply(gate, nOriginalGates);
        }
    }
    void applyMany(const std::vector<FlatMatDD>& gates, const std::vector<int>& nOriginalGates) override {
        for (auto* s : sinks_) {
            s->applyMany(gates, nOriginalGates);
        }
    }
    // the first sink owns the state
    void getState(double* real, double* imag) override { sinks_.front()->getState(real, imag); }
    void synchronize() override {
        for (auto* s : sinks_) {
            s->synchronize();
        }
    }
    void exchange(int globalPhysicalBit, int localPhysicalBit) override {
        for (auto* s : sinks_) {
            s->exchange(globalPhysicalBit, localPhysicalBit);
        }
    }
    void applyManyThenExchange(const std::vector<FlatMatDD>& gates, const std::vector<int>& nOriginalGates, int globalPhysicalBit,
                               int localPhysicalBit) override {
        for (auto* s : sinks_) {
            s->applyManyThenExchange(gates, nOriginalGates, globalPhysicalBit, localPhysicalBit);
        }
    }
    void relabel(int a, int b) override {
        for (auto* s : sinks_) {
            s->relabel(a, b);
        }
    }
    void canonicalize() override {
        for (auto* s : sinks_) {
            s->canonicalize();
        }
    }

private:
    std::vector<ArrayBackend*> sinks_;
};

} // namespace fddb200
