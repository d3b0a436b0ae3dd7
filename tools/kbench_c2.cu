// kbench_c2.cu -- developer micro-benchmark: the ACDC-sized launches of the c2 / c1 / c3 steps at their product shapes and a few
// neighbours, with the per-CTA trace (KB_DUMP=1: end time, tiles, pool draws and SM of every CTA).  Compiles in a fraction of
// kbench_tile's time (that file instantiates every sweep).   usage: kbench_c2 [reps] [only] [group]
#define KB_NO_MAIN 1
#include "kbench_tile.cu"

int main(int argc, char** argv) {
    int reps = argc > 1 ? atoi(argv[1]) : 200;
    if (argc > 2) g_only = atoi(argv[2]);
    int group = argc > 3 ? atoi(argv[3]) : 0;
    if (const char* e = getenv("DCT_TILE_REFILL")) g_refill = atoi(e);
    if (const char* e = getenv("DCT_TILE_POOL_DIV")) g_pool_div = atoi(e);
    const int64_t HW = 65536;
    using JD = JsdOp<3, true, kFwdBwd, true>;
    if (group == 0) {   // c2: B = 32, 256 x 256
        run_auto<JD, 4, 2, 4, 3, true>("jsd+dice c2", 32, HW, reps, 104, true);
        run_auto<KlLogit<true>, 4, 2, 4, 3, true>("kllogit c2", 32, HW, reps, 64, false);
        run_auto<KlFromLogits, 4, 2, 4, 3, true>("klfromlogits c2", 32, HW, reps, 48, false);
        run_auto<CopyOp<3>, 4, 2, 4, 3, true>("copy3x4", 32, HW, reps, 96, false);
    }
    if (group == 2) {   // c2 launches: fewer stages than fit (a shorter committed look-ahead: smaller end spread vs less bandwidth in flight)
        if (g_only < 0 || g_only == g_idx) run<JD, 4, 2, 4, 5, 3, true>("jsd+dice c2", 32, HW, reps, 104, true); ++g_idx;
        if (g_only < 0 || g_only == g_idx) run<JD, 4, 2, 4, 4, 3, true>("jsd+dice c2", 32, HW, reps, 104, true); ++g_idx;
        if (g_only < 0 || g_only == g_idx) run<JD, 4, 2, 4, 3, 3, true>("jsd+dice c2", 32, HW, reps, 104, true); ++g_idx;
        if (g_only < 0 || g_only == g_idx) run<KlLogit<true>, 4, 2, 4, 8, 3, true>("kllogit c2", 32, HW, reps, 64, false); ++g_idx;
        if (g_only < 0 || g_only == g_idx) run<KlLogit<true>, 4, 2, 4, 6, 3, true>("kllogit c2", 32, HW, reps, 64, false); ++g_idx;
        if (g_only < 0 || g_only == g_idx) run<KlLogit<true>, 4, 2, 4, 5, 3, true>("kllogit c2", 32, HW, reps, 64, false); ++g_idx;
        if (g_only < 0 || g_only == g_idx) run<KlLogit<true>, 4, 2, 4, 4, 3, true>("kllogit c2", 32, HW, reps, 64, false); ++g_idx;
        if (g_only < 0 || g_only == g_idx) run<KlFromLogits, 4, 2, 4, 8, 3, true>("klfromlogits c2", 32, HW, reps, 48, false); ++g_idx;
        if (g_only < 0 || g_only == g_idx) run<KlFromLogits, 4, 2, 4, 6, 3, true>("klfromlogits c2", 32, HW, reps, 48, false); ++g_idx;
        if (g_only < 0 || g_only == g_idx) run<KlFromLogits, 4, 2, 4, 4, 3, true>("klfromlogits c2", 32, HW, reps, 48, false); ++g_idx;
    }
    if (group == 3) {   // c2 launches: smaller tiles (128 / 192 pixels) for a finer end-of-launch granularity
        run_auto<JD, 4, 2, 4, 3, true>("jsd+dice c2", 32, HW, reps, 104, true);
        run_auto<JD, 4, 2, 2, 6, true>("jsd+dice c2", 32, HW, reps, 104, true);
        run_auto<JD, 4, 2, 2, 5, true>("jsd+dice c2", 32, HW, reps, 104, true);
        run_auto<JD, 4, 2, 3, 4, true>("jsd+dice c2", 32, HW, reps, 104, true);
        run_auto<JD, 4, 2, 3, 3, true>("jsd+dice c2", 32, HW, reps, 104, true);
        run_auto<KlLogit<true>, 4, 2, 2, 6, true>("kllogit c2", 32, HW, reps, 64, false);
        run_auto<KlLogit<true>, 4, 2, 3, 4, true>("kllogit c2", 32, HW, reps, 64, false);
        run_auto<KlFromLogits, 4, 2, 2, 6, true>("klfromlogits c2", 32, HW, reps, 48, false);
        run_auto<KlFromLogits, 4, 2, 3, 4, true>("klfromlogits c2", 32, HW, reps, 48, false);
    }
    if (group == 1) {   // c3: K = 2, C = 2, B = 4, 512 x 512; c1: K = 2, C = 4, B = 4, 256 x 256
        using J3 = JsdOp<2, true, kFwdBwd, true>;
        run_auto<J3, 2, 4, 8, 2>("jsd+dice c3", 4, 262144, reps, 40, true);
        run_auto<J3, 2, 2, 8, 2>("jsd+dice c3", 4, 262144, reps, 40, true);
        run_auto<J3, 2, 2, 4, 4>("jsd+dice c3", 4, 262144, reps, 40, true);
        run_auto<J3, 2, 2, 4, 4, true>("jsd+dice c3", 4, 262144, reps, 40, true);
        run_auto<J3, 2, 2, 4, 3, true>("jsd+dice c3", 4, 262144, reps, 40, true);
        run_auto<J3, 2, 4, 2, 4>("jsd+dice c3", 4, 262144, reps, 40, true);
        run_auto<J3, 2, 1, 8, 2, true>("jsd+dice c3", 4, 262144, reps, 40, true);
        run_auto<J3, 2, 1, 4, 4, true>("jsd+dice c3", 4, 262144, reps, 40, true);
        run_auto<J3, 4, 2, 4, 3, true>("jsd+dice c1", 4, HW, reps, 72, true);
        run_auto<J3, 4, 2, 4, 4, true>("jsd+dice c1", 4, HW, reps, 72, true);
        run_auto<J3, 4, 1, 4, 4, true>("jsd+dice c1", 4, HW, reps, 72, true);
        run_auto<J3, 4, 2, 2, 4, true>("jsd+dice c1", 4, HW, reps, 72, true);
        run_auto<KlLogit<true>, 2, 2, 4, 3, true>("kllogit c3", 4, 262144, reps, 32, false);
        run_auto<KlLogit<true>, 2, 2, 4, 4, true>("kllogit c3", 4, 262144, reps, 32, false);
        run_auto<KlLogit<true>, 2, 2, 8, 2>("kllogit c3", 4, 262144, reps, 32, false);
        run_auto<KlLogit<true>, 2, 4, 8, 2>("kllogit c3", 4, 262144, reps, 32, false);
    }
    return 0;
}
