// flatdd_gpu_standalone — the same command line as flatdd_gpu (and the reference's apps/FlatDD.cpp:20-146),
// built WITHOUT a reference checkout: own OpenQASM 2 reader, own gate matrices, dense-block fusion and
// flat gate-DD builder (flatdd_b200/host/standalone.hpp).  The state
// starts flat on the device, every (fused) gate is a DMAVM launch through the C-ABI.
//
//   flatdd_gpu_standalone --file C.qasm [--fuse 0|1|2|3] [--max-block 5] [--max-nondiag 4] [--gpu D]
//                         [--bin FILE] [--pv] [--shots N --seed S] [--time-gates] [--quiet]
//                         [--trace FILE [--trace-only]] [--load STATE.bin] [--world N (with --trace-only: schedule for N shards)]
// --fuse 0: one launch per gate; 1: dense-block fusion with commuting open blocks; 2: dependency-graph dense-block fusion
//         (dense matrices, GPU cost model); 3: dependency-graph fusion on block tables (<= 4 non-diagonal + <= 5 context qubits)
// (fewest launches).  Flags of the reference that only steer its DD phase (-t, --thresh, --beta, --no_cache, --DDSIM_convert,
// --ps) are accepted and ignored.  --load resumes from a state written by --bin (checkpoint / resume).
// --trace-only records the boundary traffic without touching a GPU
// (used by the CPU tests: the trace is replayed on the oracle).
#include "standalone.hpp"

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <string>

namespace {

struct Args {
    std::map<std::string, std::string> kv;
    [[nodiscard]] bool has(const std::string& k) const { return kv.count(k) > 0; }
    [[nodiscard]] std::string str(const std::string& k, const std::string& dflt = "") const {
        const auto it = kv.find(k);
        return it == kv.end() ? dflt : it->second;
    }
    [[nodiscard]] long num(const std::string& k, long dflt) const { return has(k) ? std::stol(kv.at(k)) : dflt; }
};

Args parseArgs(int argc, char** argv) {
    static const std::map<std::string, bool> takesValue = {
        {"file", true}, {"fuse", true}, {"t", true}, {"beta", true}, {"thresh", true}, {"gpu", true}, {"bin", true}, {"trace", true},
        {"shots", true}, {"seed", true}, {"load", true}, {"max-block", true}, {"max-nondiag", true}, {"budget", true}, {"world", true}, {"pv", false}, {"ps", false}, {"no_cache", false},
        {"DDSIM_convert", false}, {"trace-only", false}, {"time-gates", false}, {"quiet", false}, {"help", false}, {"h", false}};
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string k = argv[i];
        while (!k.empty() && k[0] == '-') k.erase(0, 1);
        std::string inlineValue;
        const auto eq = k.find('=');
        if (eq != std::string::npos) {
            inlineValue = k.substr(eq + 1);
            k = k.substr(0, eq);
        }
        const auto it = takesValue.find(k);
        if (it == takesValue.end()) throw std::runtime_error("unknown option --" + k);
        if (!it->second) {
            a.kv[k] = "1";
        } else if (eq != std::string::npos) {
            a.kv[k] = inlineValue;
        } else {
            if (i + 1 >= argc) throw std::runtime_error("option --" + k + " needs a value");
            a.kv[k] = argv[++i];
        }
    }
    return a;
}

} // namespace

int main(int argc, char** argv) {
    namespace sa = fddb200::standalone;
    Args args;
    try {
        args = parseArgs(argc, argv);
    } catch (const std::exception& e) {
        std::cerr << "flatdd_gpu_standalone: " << e.what() << "\n";
        return 1;
    }
    if (args.has("help") || args.has("h") || !args.has("file")) {
        std::cout << "usage: flatdd_gpu_standalone --file C.qasm [--fuse 0|1|2|3] [--max-block 5] [--max-nondiag 4] [--gpu D] [--bin FILE] [--pv]\n"
                     "                             [--shots N --seed S] [--time-gates] [--quiet] [--trace FILE [--trace-only]]\n";
        return args.has("file") ? 0 : 1;
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    std::unique_ptr<fddb200::GpuArrayBackend> gpu;
    std::unique_ptr<fddb200::TraceRecorder> recorder;
    std::unique_ptr<fddb200::TeeBackend> tee;
    std::unique_ptr<fddb200::BatchingBackend> batching;
    try {
        sa::Circuit circuit = sa::parseQasmFile(args.str("file"));
        const int nQubits = circuit.nQubits;
        const std::size_t nOps = circuit.nOps();
        const std::string name = circuit.name;
        fddb200::ArrayBackend* backend = nullptr;
        const int world = static_cast<int>(args.num("world", 1));
        if (world > 1) {
            // scheduling for a sharded state: qubit remaps and half-shard exchanges go into the trace; the multi-GPU run of a
            // trace is bench.py / flatdd_b200.sharded.replay (one process per GPU), the multi-process CLI is flatdd_gpu
            if (!args.has("trace-only") || !args.has("trace")) throw std::runtime_error("--world N schedules for N shards and needs --trace FILE --trace-only");
            if ((world & (world - 1)) != 0 || world > 8) throw std::runtime_error("--world must be 2, 4 or 8");
        }
        if (args.has("trace")) recorder = std::make_unique<fddb200::TraceRecorder>(args.str("trace"), nQubits);
        if (args.has("trace-only")) {
            if (!recorder) throw std::runtime_error("--trace-only needs --trace FILE");
            backend = recorder.get();
        } else {
            gpu = std::make_unique<fddb200::GpuArrayBackend>(nQubits, static_cast<int>(args.num("gpu", 0))); // throws without a device: no CPU fallback
            if (args.has("time-gates")) gpu->setTiming(true);
            backend = gpu.get();
            if (recorder) {
                tee = std::make_unique<fddb200::TeeBackend>(std::vector<fddb200::ArrayBackend*>{gpu.get(), recorder.get()});
                backend = tee.get();
            }
            if (!args.has("time-gates")) {
                // consecutive launches cross the boundary together: the library keeps the state tile-resident across dense blocks
                batching = std::make_unique<fddb200::BatchingBackend>(backend);
                backend = batching.get();
            }
        }
        if (world > 1) recorder->setWorldSize(world);
        sa::FlatStartSimulator sim(std::move(circuit), backend);
        sim.worldSize = world;
        sim.fuse = static_cast<unsigned>(args.num("fuse", 0));
        sim.verbose = !args.has("quiet");
        sim.policy.maxBlockQubits = static_cast<int>(args.num("max-block", sim.policy.maxBlockQubits));
        sim.policy.maxNonDiagonal = static_cast<int>(args.num("max-nondiag", sim.policy.maxNonDiagonal));
        if (args.has("budget")) sim.policy.budgetFactor = std::stod(args.str("budget"));
        if (args.has("load")) {
            // resume: the initial state is a dump written by --bin (raw little-endian fp64: real array, then imag array)
            if (!gpu) throw std::runtime_error("--load needs a GPU run");
            const std::size_t dim = std::size_t{1} << nQubits;
            std::vector<double> re(dim), im(dim);
            std::ifstream in(args.str("load"), std::ios::binary);
            if (!in.read(reinterpret_cast<char*>(re.data()), static_cast<std::streamsize>(dim * sizeof(double))) ||
                !in.read(reinterpret_cast<char*>(im.data()), static_cast<std::streamsize>(dim * sizeof(double)))) {
                throw std::runtime_error("--load: " + args.str("load") + " does not hold 2 x 2^n doubles");
            }
            fddb200::fddCheck(fdd_set_state(gpu->ctx(), re.data(), im.data()), "fdd_set_state");
            sim.stateLoaded = true;
        }
        if (sim.verbose) std::cout << (sim.stateLoaded ? "Resuming from a dumped state" : "Starting flat on the device (no DD phase)") << std::endl;
        sim.simulate();
        const std::chrono::duration<float> durationSimulation = std::chrono::high_resolution_clock::now() - t1;
        std::cout << "Simulation finished" << std::endl;
        if (recorder) recorder->close();

        if (gpu && (args.has("pv") || args.has("bin"))) {
            std::vector<double> re, im;
            sim.getVector(re, im);
            if (args.has("pv")) {
                std::ofstream out("../../log/results/state/" + name + "_FlatDD.txt");
                if (out.is_open()) {
                    for (std::size_t q = 0; q < re.size(); ++q) out << re[q] << " " << im[q] << std::endl;
                    std::cout << "Data saved to file." << std::endl;
                } else {
                    std::cerr << "Failed to open the file." << std::endl;
                }
            }
            if (args.has("bin")) {
                std::ofstream out(args.str("bin"), std::ios::binary);
                out.write(reinterpret_cast<const char*>(re.data()), static_cast<std::streamsize>(re.size() * sizeof(double)));
                out.write(reinterpret_cast<const char*>(im.data()), static_cast<std::streamsize>(im.size() * sizeof(double)));
            }
        }
        // measurement sampling on the device: most frequent outcomes
        std::string sampled;
        const auto shots = static_cast<unsigned long>(args.num("shots", 0));
        if (gpu && shots > 0) {
            std::vector<uint64_t> outcomes(shots);
            fddb200::fddCheck(fdd_sample(gpu->ctx(), shots, static_cast<uint64_t>(args.num("seed", 0)), outcomes.data()), "fdd_sample");
            std::map<uint64_t, std::size_t> histogram;
            for (uint64_t o : outcomes) ++histogram[o];
            std::vector<std::pair<std::size_t, uint64_t>> byCount;
            for (const auto& kv : histogram) byCount.emplace_back(kv.second, kv.first);
            std::sort(byCount.rbegin(), byCount.rend());
            for (std::size_t k = 0; k < byCount.size() && k < 16; ++k) {
                std::string bits(static_cast<std::size_t>(nQubits), '0');
                for (int q = 0; q < nQubits; ++q) {
                    if ((byCount[k].second >> q) & 1U) bits[static_cast<std::size_t>(nQubits - 1 - q)] = '1';
                }
                sampled += (sampled.empty() ? "" : ",\n") + std::string("    \"") + bits + "\": " + std::to_string(byCount[k].first);
            }
        }
        std::ofstream timeFile("../../log/results/time/" + name + "_FlatDD.txt");
        if (timeFile.is_open()) {
            timeFile << "Switch Overhead:" << 0 << std::endl; // no DD phase, no conversion of a grown DD
            for (const auto& t : sim.timeRecord2) timeFile << t << std::endl;
            std::cout << "Time data saved to file." << std::endl;
        }
        const double amps = std::ldexp(1.0, nQubits);
        std::printf("{\n");
        if (!sampled.empty()) std::printf("  \"samples_top16\": {\n%s\n  },\n", sampled.c_str());
        std::printf("  \"statistics\": {\n");
        std::printf("    \"DD->Array conversion\": 0.0,\n");
        std::printf("    \"applied_gates\": %zu,\n", nOps);
        std::printf("    \"array_phase_gates\": %zu,\n", sim.unitaryOps);
        std::printf("    \"array_phase_launches\": %zu,\n", sim.launches);
        std::printf("    \"array_phase_time\": %.9g,\n", sim.arrayPhaseTime);
        std::printf("    \"benchmark\": \"%s\",\n", name.c_str());
        std::printf("    \"dmavm_hbm_gbs\": %.6g,\n", sim.kernelMsTotal > 0 ? 32.0 * amps * static_cast<double>(sim.launches) / (sim.kernelMsTotal * 1e-3) / 1e9 : 0.0);
        std::printf("    \"dmavm_kernel_ms_total\": %.9g,\n", sim.kernelMsTotal);
        std::printf("    \"exchanges\": %zu,\n", sim.exchanges);
        std::printf("    \"front_end\": \"standalone (flat start, dense-block fusion)\",\n");
        std::printf("    \"gate_merging_time\": %.9g,\n", sim.gateMergingTime);
        std::printf("    \"gates_per_sec_array_phase\": %.9g,\n", sim.arrayPhaseTime > 0 ? static_cast<double>(sim.unitaryOps) / sim.arrayPhaseTime : 0.0);
        std::printf("    \"gpu_kernel_launches\": %llu,\n", gpu ? static_cast<unsigned long long>(fdd_launch_count(gpu->ctx())) : 0ULL);
        std::printf("    \"n_qubits\": %d,\n", nQubits);
        std::printf("    \"number of threads\": %ld,\n", args.num("t", 16));
        std::printf("    \"simulation_time\": %.9g,\n", static_cast<double>(durationSimulation.count()));
        std::printf("    \"switched\": true,\n");
        std::printf("    \"switched_at_op\": 0,\n");
        std::printf("    \"unitary_gates\": %zu\n", sim.unitaryOps);
        std::printf("  }\n}\n");
    } catch (const std::exception& e) {
        std::cerr << "flatdd_gpu_standalone: " << e.what() << "\n";
        return 3;
    }
    return 0;
}
