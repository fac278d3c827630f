// ref_dump.cpp — golden-vector generator.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Links the UNMODIFIED reference (oracle/_ref/libflatdd_ref.a, compiled from /root/reference by
// oracle/Makefile) and writes, for one circuit:
//   final_*.f64      the reference's own final state (SwitchSimulator::simulate(), raw fp64), or
//   samples.bin      sampled amplitudes of it when the state is too big to keep,
//   trace.bin        every flat table that crosses the drop-in boundary when the product's host
//                    driver (flatdd_b200/host/gpu_switch_simulator.hpp) runs the same circuit,
//   kat_*            known-answer tests produced by calling the reference's own functions:
//                    getVectorFromDDSwitch1 / getValueByPathPar on the DD at the switch point,
//                    DDArrMultiplyIP / DDArrMultiplyOP on seeded random states, and the cost
//                    functions DMAVMACStatIP / DMAVMACStatOP1 / size(),
//   manifest.json    what was written and the reference's statistics.
// All multi-byte values are little endian; states are SoA (real array, then imag array).
#include "SwitchSimulator.hpp"
#include "reference_binding.hpp"

#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <sys/stat.h>

namespace {

struct Args {
    std::string file, out;
    unsigned threads = 8, fuse = 0;
    double thresh = 2.0;
    bool noCache = false, fullState = false, noKat = false, noRef = false;
    int samples = 4096, katGates = 4;
    uint64_t seed = 20241017ULL;
    int traceFuse = -1; // fuse mode for the product trace (default: same as --fuse)
    int world = 1;      // shard count the product trace is scheduled for (adds exchange records)
};

[[noreturn]] void usage() {
    std::fprintf(stderr, "usage: ref_dump --file C.qasm --out DIR [-t T] [--fuse F] [--trace-fuse F] [--thresh X] [--no_cache]\n"
                         "                [--full-state] [--samples N] [--kat-gates K] [--no-kat] [--no-ref] [--seed S]\n");
    std::exit(2);
}

Args parse(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        const std::string k = argv[i];
        auto val = [&]() -> std::string {
            if (i + 1 >= argc) usage();
            return argv[++i];
        };
        if (k == "--file") a.file = val();
        else if (k == "--out") a.out = val();
        else if (k == "-t") a.threads = static_cast<unsigned>(std::stoul(val()));
        else if (k == "--fuse") a.fuse = static_cast<unsigned>(std::stoul(val()));
        else if (k == "--trace-fuse") a.traceFuse = std::stoi(val());
        else if (k == "--thresh") a.thresh = std::stod(val());
        else if (k == "--no_cache") a.noCache = true;
        else if (k == "--full-state") a.fullState = true;
        else if (k == "--no-kat") a.noKat = true;
        else if (k == "--no-ref") a.noRef = true;
        else if (k == "--samples") a.samples = std::stoi(val());
        else if (k == "--kat-gates") a.katGates = std::stoi(val());
        else if (k == "--seed") a.seed = std::stoull(val());
        else if (k == "--world") a.world = std::stoi(val());
        else usage();
    }
    if (a.file.empty() || a.out.empty()) usage();
    if (a.traceFuse < 0) a.traceFuse = static_cast<int>(a.fuse);
    return a;
}

void writeRaw(const std::string& path, const void* p, std::size_t bytes) {
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (f == nullptr || (bytes != 0 && std::fwrite(p, 1, bytes, f) != bytes)) {
        throw std::runtime_error("cannot write " + path);
    }
    std::fclose(f);
}

template <int R> void writeFlat(const std::string& path, const fddb200::FlatDD<R>& dd) {
    // int32 n_qubits, n_nodes, root, radix; double root_weight[2]; level; child; weight
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (f == nullptr) throw std::runtime_error("cannot write " + path);
    const int32_t head[4] = {dd.n_qubits, dd.nNodes(), dd.root, R};
    std::fwrite(head, sizeof head, 1, f);
    std::fwrite(dd.root_weight, sizeof dd.root_weight, 1, f);
    std::fwrite(dd.level.data(), sizeof(int32_t), dd.level.size(), f);
    std::fwrite(dd.child.data(), sizeof(int32_t), dd.child.size(), f);
    std::fwrite(dd.weight.data(), sizeof(double), dd.weight.size(), f);
    std::fclose(f);
}

using RefSim = SwitchSimulator<dd::DDPackageConfig>;

std::string pathBits(std::size_t i, std::size_t n) { // LSB first, as getValueByPath indexes elements.at(v)
    std::string s(n, '0');
    for (std::size_t q = 0; q < n; ++q) {
        if ((i >> q) & 1U) s[q] = '1';
    }
    return s;
}

} // namespace

int main(int argc, char** argv) {
    const Args a = parse(argc, argv);
    ::mkdir(a.out.c_str(), 0755);
    std::ostringstream man;
    man.precision(17);
    man << "{\n  \"circuit\": \"" << a.file.substr(a.file.find_last_of('/') + 1) << "\",\n";
    man << "  \"threads\": " << a.threads << ", \"fuse\": " << a.fuse << ", \"trace_fuse\": " << a.traceFuse
        << ", \"thresh\": " << a.thresh << ", \"no_cache\": " << (a.noCache ? "true" : "false") << ",\n";

    const unsigned nThreadExp = static_cast<unsigned>(std::log2(a.threads));

    // ---- 1. the product's host driver, recording the boundary traffic ---------------------
    std::size_t nQubits = 0;
    {
        auto qcp = std::make_unique<qc::QuantumComputation>(a.file);
        nQubits = qcp->getNqubits();
        fddb200::TraceRecorder rec(a.out + "/trace.bin", static_cast<int>(nQubits));
        if (a.world > 1) rec.setWorldSize(a.world);
        fddb200::RefGpuSwitchSimulator sim(std::move(qcp), &rec);
        sim.worldSize = a.world;
        sim.threshold = a.thresh;
        sim.n_thread_exp = nThreadExp;
        sim.fuse = static_cast<unsigned>(a.traceFuse);
        sim.enable_cache = !a.noCache;
        sim.verbose = false;
        sim.simulate();
        if (!sim.switched) {
            // never switched: the CLI's --pv path converts the final DD (apps/FlatDD.cpp:89-93)
            sim.getVectorFromDD();
        }
        rec.close();
        man << "  \"n_qubits\": " << nQubits << ", \"n_ops\": " << sim.getNumberOfOps() << ",\n";
        man << "  \"trace\": {\"records\": " << rec.records() << ", \"switched\": " << (sim.switched ? "true" : "false")
            << ", \"switched_at_op\": " << sim.switchedAtOp << ", \"unitary_ops\": " << sim.unitaryOps
            << ", \"array_phase_ops\": " << sim.arrayPhaseOps << ", \"launches\": " << sim.launches
            << ", \"world\": " << a.world << ", \"exchanges\": " << sim.exchanges
            << ", \"gate_merging_s\": " << sim.gateMergingTime << "},\n";
    }
    const std::size_t dim = std::size_t{1} << nQubits;

    if (a.noRef) {
        man << "  \"reference\": null\n}\n";
        writeRaw(a.out + "/manifest.json", man.str().data(), man.str().size());
        return 0;
    }

    // ---- 2. the reference itself --------------------------------------------------------------
    auto qcp = std::make_unique<qc::QuantumComputation>(a.file);
    auto ref = std::make_unique<RefSim>(std::move(qcp));
    ref->threshold = a.thresh;
    ref->n_thread_exp = nThreadExp;
    ref->fuse = a.fuse;
    ref->enable_cache = !a.noCache;
    ref->switchTime = 0.0;
    const auto t0 = std::chrono::steady_clock::now();
    ref->simulate();
    const double simSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int finalIdx = 0;
    if (ref->switched) {
        finalIdx = ref->stateVecIdx;
    } else {
        ref->getVectorFromDD(0); // what apps/FlatDD.cpp:89-93 does for --pv
    }
    const double* fr = ref->state_real[static_cast<std::size_t>(finalIdx)];
    const double* fi = ref->state_imag[static_cast<std::size_t>(finalIdx)];
    double norm2 = 0.0;
    for (std::size_t i = 0; i < dim; ++i) norm2 += fr[i] * fr[i] + fi[i] * fi[i];
    double array_s = 0.0;
    for (double t : ref->timeRecord2) array_s += t;
    man << "  \"reference\": {\"switched\": " << (ref->switched ? "true" : "false") << ", \"simulate_s\": " << simSeconds
        << ", \"switch_s\": " << ref->switchTime << ", \"array_phase_s\": " << array_s
        << ", \"array_phase_launches\": " << ref->timeRecord2.size() << ", \"dd_phase_steps\": " << ref->timeRecord1.size()
        << ", \"norm2\": " << norm2 << ", \"final_dd_size\": " << ref->dd->size(ref->rootEdge) << "},\n";

    if (a.fullState) {
        writeRaw(a.out + "/final_re.f64", fr, dim * sizeof(double));
        writeRaw(a.out + "/final_im.f64", fi, dim * sizeof(double));
    }
    if (!a.fullState) {
        // sampled amplitudes: seeded uniform indices plus the 256 largest |amp|^2
        std::mt19937_64 rng(a.seed);
        std::vector<uint64_t> idx;
        const int nS = static_cast<int>(std::min<std::size_t>(static_cast<std::size_t>(a.samples), dim));
        for (int s = 0; s < nS; ++s) idx.push_back(rng() % dim);
        std::vector<std::pair<double, uint64_t>> top;
        for (std::size_t i = 0; i < dim; ++i) {
            const double p = fr[i] * fr[i] + fi[i] * fi[i];
            if (top.size() < 256) {
                top.emplace_back(p, i);
                std::push_heap(top.begin(), top.end(), std::greater<>());
            } else if (p > top.front().first) {
                std::pop_heap(top.begin(), top.end(), std::greater<>());
                top.back() = {p, i};
                std::push_heap(top.begin(), top.end(), std::greater<>());
            }
        }
        for (auto& t : top) idx.push_back(t.second);
        std::FILE* f = std::fopen((a.out + "/samples.bin").c_str(), "wb");
        const uint64_t cnt = idx.size();
        std::fwrite(&cnt, sizeof cnt, 1, f);
        std::fwrite(idx.data(), sizeof(uint64_t), idx.size(), f);
        for (uint64_t i : idx) std::fwrite(&fr[i], sizeof(double), 1, f);
        for (uint64_t i : idx) std::fwrite(&fi[i], sizeof(double), 1, f);
        std::fclose(f);
    }

    // ---- 3. known-answer tests from the reference's own functions -----------------------------
    man << "  \"kats\": [";
    bool firstKat = true;
    auto katSep = [&]() {
        if (!firstKat) man << ",";
        firstKat = false;
        man << "\n    ";
    };
    if (!a.noKat) {
        auto& pkg = ref->dd;
        pkg->n_thread_exp = nThreadExp;
        // 3a. conversion of the DD the reference holds (at the switch point, or the final DD)
        {
            const auto flat = fddb200::flattenVector(ref->rootEdge, static_cast<int>(nQubits));
            writeFlat(a.out + "/kat_convert_dd.bin", flat);
            const int scratch = 1 - finalIdx;
            std::vector<double> keepR(fr, fr + dim), keepI(fi, fi + dim);
            auto* sr = ref->state_real[static_cast<std::size_t>(scratch)];
            auto* si = ref->state_imag[static_cast<std::size_t>(scratch)];
            std::memset(sr, 0, dim * sizeof(double));
            std::memset(si, 0, dim * sizeof(double));
            ref->getVectorFromDDSwitch1(scratch);
            writeRaw(a.out + "/kat_convert_switch1_re.f64", sr, dim * sizeof(double));
            writeRaw(a.out + "/kat_convert_switch1_im.f64", si, dim * sizeof(double));
            std::vector<double> wr(dim), wi(dim);
            for (std::size_t i = 0; i < dim; ++i) {
                const auto cv = pkg->getValueByPathPar(ref->rootEdge, pathBits(i, nQubits));
                wr[i] = cv.r;
                wi[i] = cv.i;
            }
            writeRaw(a.out + "/kat_convert_walk_re.f64", wr.data(), dim * sizeof(double));
            writeRaw(a.out + "/kat_convert_walk_im.f64", wi.data(), dim * sizeof(double));
            std::memset(sr, 0, dim * sizeof(double));
            std::memset(si, 0, dim * sizeof(double));
            katSep();
            man << "{\"kind\": \"convert\", \"dd\": \"kat_convert_dd.bin\", \"nodes\": " << flat.nNodes()
                << ", \"dd_size\": " << pkg->size(ref->rootEdge) << ", \"threads\": " << a.threads << "}";
        }
        // 3b. DMAVM on seeded random states: single gates spread over the circuit and 6-gate products
        {
            // fresh parse: the reference simulator owns (and hides) its circuit
            const auto circuit = std::make_unique<qc::QuantumComputation>(a.file);
            std::vector<const qc::Operation*> unitary;
            for (auto& op : *circuit) {
                if (!op->isNonUnitaryOperation() && !op->isClassicControlledOperation()) unitary.push_back(op.get());
            }
            std::mt19937_64 rng(a.seed + 1);
            std::normal_distribution<double> gauss(0.0, 1.0);
            std::vector<double> yr(dim), yi(dim), zr(dim), zi(dim), zor(dim), zoi(dim);
            std::vector<double*> scratchR(a.threads), scratchI(a.threads);
            for (unsigned t = 0; t < a.threads; ++t) {
                scratchR[t] = static_cast<double*>(std::calloc(dim, sizeof(double)));
                scratchI[t] = static_cast<double*>(std::calloc(dim, sizeof(double)));
            }
            const int nK = std::min<int>(a.katGates, static_cast<int>(unitary.size()));
            for (int k = 0; k < 2 * nK; ++k) {
                const bool fused = k >= nK;
                const std::size_t at = unitary.size() * static_cast<std::size_t>(fused ? k - nK : k) / static_cast<std::size_t>(nK);
                dd::mEdge gate = dd::getDD(unitary[at], pkg);
                int count = 1;
                if (fused) {
                    for (std::size_t j = at + 1; j < unitary.size() && count < 6; ++j, ++count) {
                        gate = pkg->multiply(dd::getDD(unitary[j], pkg), gate);
                    }
                }
                double nrm = 0.0;
                for (std::size_t i = 0; i < dim; ++i) {
                    yr[i] = gauss(rng);
                    yi[i] = gauss(rng);
                    nrm += yr[i] * yr[i] + yi[i] * yi[i];
                }
                nrm = 1.0 / std::sqrt(nrm);
                for (std::size_t i = 0; i < dim; ++i) {
                    yr[i] *= nrm;
                    yi[i] *= nrm;
                }
                std::fill(zr.begin(), zr.end(), 0.0);
                std::fill(zi.begin(), zi.end(), 0.0);
                pkg->DDArrMultiplyIP(gate, yr.data(), yi.data(), zr.data(), zi.data(), dim);
                std::fill(zor.begin(), zor.end(), 0.0);
                std::fill(zoi.begin(), zoi.end(), 0.0);
                pkg->DDArrMultiplyOP(gate, yr.data(), yi.data(), zor.data(), zoi.data(), dim, scratchR, scratchI);
                double maxDiffOp = 0.0;
                for (std::size_t i = 0; i < dim; ++i) {
                    maxDiffOp = std::max(maxDiffOp, std::max(std::abs(zr[i] - zor[i]), std::abs(zi[i] - zoi[i])));
                }
                std::unordered_map<dd::mNode*, std::size_t> macMap;
                const std::size_t costIp = pkg->DMAVMACStatIP(gate, macMap, dim, nThreadExp);
                macMap.clear();
                const std::size_t costOp1 = pkg->DMAVMACStatOP1(gate, macMap, dim, nThreadExp);
                macMap.clear();
                const std::size_t nnz = pkg->DMAVMACStatIP(gate, macMap, dim, 0);
                const auto flat = fddb200::flattenMatrix(gate, static_cast<int>(nQubits));
                const std::string stem = "kat_gate" + std::to_string(k);
                writeFlat(a.out + "/" + stem + "_dd.bin", flat);
                writeRaw(a.out + "/" + stem + "_y_re.f64", yr.data(), dim * sizeof(double));
                writeRaw(a.out + "/" + stem + "_y_im.f64", yi.data(), dim * sizeof(double));
                writeRaw(a.out + "/" + stem + "_z_re.f64", zr.data(), dim * sizeof(double));
                writeRaw(a.out + "/" + stem + "_z_im.f64", zi.data(), dim * sizeof(double));
                katSep();
                man << "{\"kind\": \"dmavm\", \"stem\": \"" << stem << "\", \"op_index\": " << at << ", \"fused_ops\": " << count
                    << ", \"nodes\": " << flat.nNodes() << ", \"dd_size\": " << pkg->size(gate) << ", \"nnz\": " << nnz
                    << ", \"cost_ip\": " << costIp << ", \"cost_op1\": " << costOp1 << ", \"threads\": " << a.threads
                    << ", \"max_abs_diff_op_vs_ip\": " << maxDiffOp << "}";
            }
            for (unsigned t = 0; t < a.threads; ++t) {
                std::free(scratchR[t]);
                std::free(scratchI[t]);
            }
        }
    }
    man << "\n  ]\n}\n";
    writeRaw(a.out + "/manifest.json", man.str().data(), man.str().size());
    return 0;
}
