// caf_map_cuda_harness.cpp -- TEST HARNESS for integration/blackscholes.c.caf_cuda.patch.
//
// CAF is not vendored in P3ARSEC (parsec-ff/config/gcc-caf.bldconf:11-18 points at an external install), so the CAF
// build of the reference cannot be compiled here.  This program reproduces, without CAF, exactly what the reference's
// CAF_V3 driver does around the Map, so that the handler the patch installs in map_func can be run on a GPU:
//   * struct DataCont / InVec / OutVec                          blackscholes.c:482-492
//   * the loader and the AoS->SoA loop                           blackscholes.c:696-767
//   * the message: data_vec(numOptions) then numOptions push_backs -- 2N records, the first N zero  :771-778
//   * NUM_RUNS requests, each answered with a vector of prices  blackscholes.c:885-895
// The handler body below is the one in the patch, verbatim.  The reference's CAF build never copies final_res into
// `prices` (its output file is uninitialised memory); this harness writes final_res[N..2N), the prices of the real
// records, so that the test can compare them with the golden outputs of the other variants.
//   g++ -O2 -std=c++11 -I include integration/caf_map_cuda_harness.cpp -o harness -L p3arsec_b200/lib -lbs_gpu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <bs_gpu.h>

#define fptype float
#define NUM_RUNS 100

struct DataCont {
    int otype;
    float sptprice;
    float strike;
    float rate;
    float volatility;
    float otime;
};
using InVec = std::vector<DataCont>;
using OutVec = std::vector<fptype>;

// ---- the handler of integration/blackscholes.c.caf_cuda.patch (map_func's message handler) ----
static OutVec map_func_handler(const InVec &data_vec, uint64_t nw_)
{
    static bs_gpu_ctx *gpu = NULL;
    const size_t nv = data_vec.size();
    if (gpu == NULL) {
        int nGpus = bs_gpu_device_count();
        if ((uint64_t)nGpus > nw_) nGpus = (int)nw_;
        if (nGpus < 1 || bs_gpu_init(&gpu, nGpus, nv, sizeof(fptype)) != BS_GPU_OK) {
            printf("ERROR: Unable to initialise the GPU context.\n");
            exit(1);
        }
    }
    OutVec res(nv);
    // one message = one run of the Map (the driver sends NUM_RUNS messages)
    if (bs_gpu_price_aos(gpu, data_vec.data(), nv, res.data(), 1) != BS_GPU_OK) {
        printf("ERROR: bs_gpu_price_aos failed: %s\n", bs_gpu_last_error(gpu));
        exit(1);
    }
    return res;
}

int main(int argc, char **argv)
{
    if (argc != 4) {
        printf("Usage:\n\t%s <nthreads> <inputFile> <outputFile>\n", argv[0]);
        return 1;
    }
    const int nThreads = atoi(argv[1]);
    FILE *file = fopen(argv[2], "r");
    if (!file) { printf("ERROR: Unable to open file `%s'.\n", argv[2]); return 1; }
    int numOptions = 0;
    if (fscanf(file, "%i", &numOptions) != 1) { printf("ERROR: Unable to read from file `%s'.\n", argv[2]); return 1; }
    std::vector<int> otype(numOptions);
    std::vector<fptype> sptprice(numOptions), strike(numOptions), rate(numOptions), volatility(numOptions), otime(numOptions);
    for (int i = 0; i < numOptions; i++) {
        float s, k, r, divq, v, t, divs, ref;
        char ty;
        if (fscanf(file, "%f %f %f %f %f %f %c %f %f", &s, &k, &r, &divq, &v, &t, &ty, &divs, &ref) != 9) {
            printf("ERROR: Unable to read from file `%s'.\n", argv[2]);
            return 1;
        }
        otype[i] = (ty == 'P') ? 1 : 0;
        sptprice[i] = s; strike[i] = k; rate[i] = r; volatility[i] = v; otime[i] = t;
    }
    fclose(file);

    InVec data_vec(numOptions);                       // :772 -- numOptions zero records ...
    for (int i = 0; i < numOptions; i++) {
        DataCont data{otype[i], sptprice[i], strike[i], rate[i], volatility[i], otime[i]};
        data_vec.push_back(std::move(data));          // :776 -- ... followed by the numOptions real ones
    }
    OutVec final_res(numOptions);
    const uint64_t nw = (uint64_t)(nThreads < 1 ? 1 : nThreads);
    for (uint32_t j = 0; j < NUM_RUNS; j++) final_res = map_func_handler(data_vec, nw);   // :885-895

    printf("message records: %zu, prices returned: %zu\n", data_vec.size(), final_res.size());
    FILE *out = fopen(argv[3], "w");
    if (!out) { printf("ERROR: Unable to open file `%s'.\n", argv[3]); return 1; }
    fprintf(out, "%i\n", numOptions);
    for (int i = 0; i < numOptions; i++) fprintf(out, "%.18f\n", final_res[(size_t)numOptions + i]);
    fclose(out);
    // the zero records price to NaN, on the GPU as on the CPU
    int nan_head = 0;
    for (int i = 0; i < numOptions; i++) nan_head += (final_res[i] != final_res[i]);
    printf("zero records priced NaN: %d of %d\n", nan_head, numOptions);
    return 0;
}
