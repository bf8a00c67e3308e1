/* oracle/az_oracle.h — C face of the CPU restatement of the reference's self-play hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (alphazero-pybind11_b200/) never does.
 *
 * What it restates (reference file:line in az_oracle.cc next to each function):
 *   Connect4GS            src/connect4_gs.cc:39-149
 *   Node / MCTS (PUCT)    src/mcts.cc:93-173, 403-555, 557-750; src/mcts.h:78-100
 *   PlayManager::play()   src/play_manager.cc:258-600, update_inferences :619-642,
 *                         build_batch / build_history_batch  src/py_wrapper.cc:393-424, 449-504
 *   dumb_eval             src/game_state.h:160-173
 *   pcg32                 src/pcg/pcg_random.hpp:484-501, 1663 (setseq_xsh_rr_64_32)
 * Random draws go through libstdc++'s own std::shuffle / gamma_distribution /
 * uniform_real_distribution (the same library code the reference instantiates), so this port also
 * pins the product's hand-written device RNG (csrc/az_rng.h).
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py runs this port and the unmodified
 * reference (oracle/_ref/libazref.so, built from /root/reference by oracle/Makefile) on the same
 * seeds and requires bit-identical move lists, visit counts, Q values and history targets; the
 * committed fixtures under tests/golden/ were generated from the unmodified reference
 * (tools/make_golden.py) and are checked against this port on boxes where /root/reference is absent.
 *
 * rng_mode 1 (global) = one pcg32(seed) shared by all games, consumed in the reference's
 * single-worker order; rng_mode 0 (per game) = slot g draws from pcg32(seed + g) on the default
 * stream, i.e. slot g replays what the UNMODIFIED reference PlayManager does with concurrent_games = 1
 * after MCTS::seed_thread_rng(seed + g) — the "scalable" parity level, pinned directly to the reference
 * slot by slot (tests/test_gpu_parity.py::test_fused_kernel_slots_equal_reference_single_game_runs).
 */
#ifndef AZ_ORACLE_H_
#define AZ_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct azo_cfg {
  uint32_t games_to_play, concurrent_games;
  uint32_t mcts_visits[2];
  float cpuct, start_temp, final_temp, temp_decay_half_life;
  uint8_t history_enabled, tree_reuse, root_fpu_zero, shaped_dirichlet;
  uint8_t policy_target_pruning, eval_type /* 0 NN, 1 RANDOM */, rng_mode /* 0 per game, 1 global */, pad_;
  float epsilon, mcts_root_temp, fpu_reduction;
  uint64_t seed;
  /* playout-cap randomisation and resign (play_manager.cc:306-337, 523-524, 559-560). The reference flips these
   * coins with a thread_local, unseedable std::default_random_engine; here they come from the same stream as every
   * other draw (per game or global), which is what the engine does — a parity definition of this repo, not of the
   * reference. */
  uint8_t playout_cap_randomization, pad2_[3];
  uint32_t playout_cap_depth;
  float playout_cap_percent, resign_percent, resign_playthrough_percent;
  /* Gumbel AlphaZero (play_manager.h:104-116, mcts.cc:175-401) */
  uint8_t gumbel_enabled, gumbel_full, fast_search_uses_gumbel, pad3_;
  uint32_t gumbel_m;
  float gumbel_c_visit, gumbel_c_scale;
} azo_cfg;

void* azo_pm_new(const azo_cfg* c);
void azo_pm_free(void* h);
/* Drain awaiting_mcts_ exactly like ONE reference worker thread would (FIFO). With RANDOM eval
 * this plays every game to the end; with NN eval it returns once every active game waits for an
 * evaluation. Returns the number of leaves waiting. */
uint32_t azo_pm_run(void* h);
/* Same, but at most `n` loop iterations (RANDOM eval never drains: n = G * generations while every slot
 * is active). */
uint32_t azo_pm_run_iterations(void* h, uint64_t n);
uint32_t azo_pm_build_batch(void* h, uint32_t max, uint32_t* ids, float* canon /* [max][168] */);
void azo_pm_update_inferences(void* h, const uint32_t* ids, uint32_t n, const float* v /* [n][3] */,
                              const float* pi /* [n][7] */);
uint32_t azo_pm_drain_history(void* h, uint32_t max, float* canon, float* v, float* pi);
uint32_t azo_pm_hist_count(void* h);
uint32_t azo_pm_games_completed(void* h);
uint32_t azo_pm_remaining_games(void* h);
uint64_t azo_pm_simulations(void* h);
uint64_t azo_pm_moves(void* h);
void azo_pm_scores(void* h, float* out3);
void azo_pm_resign_scores(void* h, float* out3);
/* avg_game_length, avg_leaf_depth, avg_search_entropy, fast_avg_leaf_depth, fast_avg_search_entropy,
 * avg_moves_per_turn, avg_valid_moves (play_manager.h:288-315) */
void azo_pm_metrics(void* h, float* out7);
void azo_pm_peek(void* h, uint32_t game, uint32_t seat, uint8_t* state89, uint32_t* counts7, float* q7,
                 float* root_value3, uint32_t* depth, uint32_t* root_n, float* policy7);

/* Connect4 rules on the reference's int8[2][6][7] board. */
int azo_c4_play(int8_t* board84, uint8_t* player, uint32_t* turn, uint32_t move); /* 0 ok, -1 full column */
void azo_c4_valid(const int8_t* board84, uint8_t* out7);
int azo_c4_scores(const int8_t* board84, float* out3); /* 1 if terminal */
void azo_c4_canonical(const int8_t* board84, uint8_t player, float* out168);

/* RNG pieces, for pinning csrc/az_rng.h */
void* azo_rng_new(uint64_t seed, int use_stream, uint64_t stream);
void azo_rng_free(void* r);
uint32_t azo_rng_u32(void* r);
void azo_rng_shuffle(void* r, uint32_t n, uint32_t* inout);
float azo_rng_uniform01(void* r);
void azo_rng_gamma(void* r, float alpha, uint32_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif
