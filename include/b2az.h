/* b2az.h — C ABI of the B200-native batched self-play engine (libb2az.so).
 *
 * The reference (bhansconnect/alphazero-pybind11) has no C ABI: its boundary is the pybind11
 * module `alphazero` (src/py_wrapper.cc:108-788). This header is the layer the new build adds
 * UNDERNEATH that module: plain pointers and sizes, no torch / pybind / C++ types. Each entry
 * point names the reference interface it replaces; INTEGRATION.md shows the binding a maintainer
 * adds on the reference side (pybind11 and ctypes stubs).
 *
 * Conventions: every function returns 0 on success, a negative B2AZ_E* code otherwise, and the
 * text of the last error on the calling thread is available from b2az_last_error(). C++
 * exceptions never cross this boundary. Pointers named *_dev are CUDA device pointers, *_host are
 * host pointers (pinned or pageable). `stream` is a cudaStream_t passed as void* (NULL = the
 * legacy default stream). All device work is issued on that stream; calls taking host buffers
 * synchronise it before returning.
 */
#ifndef B2AZ_H_
#define B2AZ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2AZ_OK 0
#define B2AZ_EINVAL (-1)   /* bad argument / unsupported parameter combination */
#define B2AZ_ECUDA (-2)    /* CUDA runtime error */
#define B2AZ_ENOMEM (-3)   /* node pool / history ring exhausted on the device */
#define B2AZ_ESTATE (-4)   /* call not valid in the current state (e.g. step with evals pending) */
#define B2AZ_EMOVE (-5)    /* illegal move (reference: std::runtime_error, connect4_gs.cc:57, mcts.cc:161) */

#define B2AZ_GAME_CONNECT4 0

#define B2AZ_STEP_DEFAULT 0 /* = B2AZ_STEP_SYNC */
#define B2AZ_STEP_QUEUE 4   /* persistent CTAs, game state in shared memory, work queues (az_engine_queue.h) */
#define B2AZ_STEP_FLAT 1    /* one thread per game slot, flattened loop (round-1 kernel) */
#define B2AZ_STEP_SYNC 3    /* one thread per game slot, the 32 games of a warp in lock step (run_sync) */
#define B2AZ_STEP_WAVES 2   /* persistent CTAs, game state in shared memory, barrier-separated waves (az_engine_waves.h) */

#define B2AZ_EVAL_NN 0      /* EvalType::NN      (play_manager.h:20) */
#define B2AZ_EVAL_RANDOM 1  /* EvalType::RANDOM: dumb_eval on the device (game_state.h:160-173) */

#define B2AZ_RNG_PER_GAME 0 /* one generator per game slot: slot g draws from pcg32(seed + g), i.e. it reproduces a
                               reference PlayManager with concurrent_games = 1 run after
                               MCTS::seed_thread_rng(seed + g) (mcts.cc:19-21), bit for bit */
#define B2AZ_RNG_GLOBAL 1   /* ONE pcg32(seed) consumed in ascending game order, exactly like a single
                               reference worker thread after MCTS::seed_thread_rng(seed) (mcts.cc:19-21).
                               Serial by construction; this is the bit-exact parity mode. */

/* Mirrors PlayParams (play_manager.h:60-154) for the Connect4 engine, plus engine sizing. Fields keep the
 * reference's names and meaning: PUCT and Gumbel search (incl. gumbel_full), playout-cap randomisation, resignation,
 * per-seat visit budgets, two model groups and seat permutations are carried. Not carried: the other per-seat 2-D
 * overrides (the host side folds them when they are uniform), playout eval. b2az_create rejects a non-default value
 * for anything it does not implement instead of ignoring it. */
typedef struct b2az_params {
  uint32_t game;                 /* B2AZ_GAME_* */
  uint32_t games_to_play;
  uint32_t concurrent_games;
  uint32_t max_batch_size;       /* legacy build_batch cap; 0 = concurrent_games */
  uint32_t max_cache_size;       /* entries; 0 = no cache */
  uint32_t mcts_visits[2];       /* per seat (mcts_visits, play_manager.h:82) */
  float cpuct;
  float start_temp, final_temp, temp_decay_half_life;
  uint8_t history_enabled;
  uint8_t self_play;
  uint8_t tree_reuse;
  uint8_t playout_cap_randomization;
  float epsilon;
  float mcts_root_temp;
  uint32_t playout_cap_depth;
  float playout_cap_percent;
  float fpu_reduction;
  uint8_t root_fpu_zero;
  uint8_t shaped_dirichlet;
  uint8_t policy_target_pruning;
  uint8_t gumbel_enabled;        /* Gumbel root search (mcts.cc:175-401); parameters at the end of the struct */
  float resign_percent;          /* with playout_cap_randomization: B2AZ_RNG_PER_GAME only (coins from the game's stream) */
  float resign_playthrough_percent;
  uint8_t eval_type;             /* B2AZ_EVAL_*; applies to every seat */
  uint8_t rng_mode;              /* B2AZ_RNG_* */
  uint8_t per_slot_quota;        /* 1: every slot plays exactly games_to_play / concurrent_games games and then retires
                                    (which slot plays the last games of a run otherwise depends on completion order,
                                    play_manager.cc:506-513). Slot g then equals a reference PlayManager with
                                    concurrent_games = 1, games_to_play = quota, seeded seed + g. */
  uint8_t pad1_;
  uint64_t seed;
  /* engine sizing (no reference counterpart) */
  uint64_t pool_nodes;           /* tree-node pool size in nodes (7 per 160 B child block); 0 = sized from visits and free HBM */
  uint32_t history_capacity;     /* finished-sample ring, in samples; 0 = default */
  uint32_t lanes_per_game;       /* threads per game slot: 0 or 1 (Connect4 runs one thread per game) */
  uint32_t compact_pages;        /* a tree's arena (10 KB pages of 64 blocks) is compacted at a move once it holds more
                                    pages than this; 0 = half of the tree's share of the pool */
  /* Gumbel AlphaZero (play_manager.h:104-116) */
  uint32_t gumbel_m;             /* 16 */
  float gumbel_c_visit;          /* 50 */
  float gumbel_c_scale;          /* 1 */
  uint8_t gumbel_full;           /* pi'-matching at interior nodes too */
  uint8_t fast_search_uses_gumbel;
  uint8_t model_groups[2];       /* model group (0 or 1) of each seat (PlayParams::model_groups, play_manager.cc:24-31): the
                                    position cache keeps the groups apart, b2az_leaf_seats_host tells which seat searches */
  uint32_t step_kernel;          /* B2AZ_STEP_*: which fused step kernel runs B2AZ_RNG_PER_GAME (results do not depend on it) */
  uint32_t seat_cap_visits[2];   /* per-seat fast-search budget (seat_cap_visits, play_manager.cc:82-90); 0 = playout_cap_depth */
  /* Seat permutations (PlayParams::seat_perms, play_manager.cc:46-90, 213-221; what game_runner.play_past and
   * tournament.py build): seat_perms[p][seat] = the model group (0 or 1) that searches for `seat` in the games of
   * permutation p. Slot g plays permutation g % n_seat_perms in every one of its games (the reference's round-robin
   * hand-out when the slots finish in order), so concurrent_games must be a multiple of n_seat_perms.
   * 0 = one seating given by model_groups. */
  uint32_t n_seat_perms;         /* <= 8 */
  uint8_t seat_perms[8][2];
  uint32_t perm_seat_visits[8][2];      /* seat_visits_[p][seat] (play_manager.cc:70-80); 0 = mcts_visits[seat] */
  uint32_t perm_seat_cap_visits[8][2];  /* seat_cap_visits_[p][seat] (:82-90); 0 = seat_cap_visits[seat] */
  uint8_t group_random[2];       /* eval_type NN only: model group i is EvalType::RANDOM (eval_types_[group],
                                    play_manager.cc:578-587): its searches run dumb_eval on the device and never
                                    show up in the leaf batch */
  uint8_t pad5_[2];
} b2az_params;

/* PlayManager's per-permutation tables (perm_scores_, variant_perm_scores_: play_manager.cc:205-255, 466-474). */
typedef struct b2az_perm_stats {
  float scores[3];
  uint32_t games_completed;
  float variant_scores[4][3];    /* games with variants (StarGambitUnifiedGS) */
  uint32_t variant_games_completed[4];
} b2az_perm_stats;

/* Counters and metrics of PlayManager (play_manager.h:173-366). */
typedef struct b2az_stats {
  uint64_t simulations;          /* find_leaf+process_result pairs completed (mcts.cc:553 ++depth_) */
  uint64_t moves;                /* game moves played */
  uint32_t games_completed;      /* games_completed() */
  uint32_t games_started;
  uint32_t active_games;         /* slots still cycling */
  uint32_t leaf_count;           /* leaves waiting for an evaluation (awaiting_inference_count) */
  uint32_t hist_count;           /* hist_count() */
  float scores[3];               /* scores(): wins seat0, wins seat1, draws */
  float resign_scores[3];
  float avg_game_length, avg_leaf_depth, avg_search_entropy;
  float fast_avg_leaf_depth, fast_avg_search_entropy;
  float avg_moves_per_turn, avg_valid_moves;
  uint64_t cache_hits, cache_misses, cache_evictions, cache_reinserts, cache_size, cache_max_size;
  uint64_t pool_pages_total, pool_pages_free; /* tree-node pool occupancy */
  uint32_t device_error;         /* sticky device-side error bits (B2AZ_DEVERR_*) */
  uint32_t pad_;
  uint64_t compactions;          /* tree arenas compacted (Cheney copies) so far */
  /* the raw accumulators behind the means above (play_manager.h:288-315), all over COMPLETED games: what a multi-GPU run
   * all-reduces (a mean of means would weigh the ranks wrongly) */
  uint64_t sum_game_length;      /* game_length_ */
  uint64_t total_move_count, full_move_count, fast_move_count;
  double sum_leaf_depth, sum_search_entropy;            /* over full searches */
  double fast_sum_leaf_depth, fast_sum_search_entropy;  /* over capped (fast) searches */
  double sum_valid_moves;                               /* over all moves */
} b2az_stats;

#define B2AZ_DEVERR_POOL 1u      /* node pool exhausted */
#define B2AZ_DEVERR_HIST 2u      /* history ring overflow (samples dropped) */
#define B2AZ_DEVERR_MOVE 4u      /* update_root could not find the move (mcts.cc:159-162) */
#define B2AZ_DEVERR_DEPTH 8u     /* selection path longer than the path buffer */
#define B2AZ_DEVERR_QUEUE 16u    /* the step kernel's work-queue watchdog fired */

typedef struct b2az_engine b2az_engine;

const char* b2az_last_error(void);
int b2az_params_default(b2az_params* p);

/* PlayManager::PlayManager(gs, params) (play_manager.cc:12-256): allocates the device pool on
 * `device`, initialises concurrent_games slots (fresh game, one tree per seat). */
int b2az_create(const b2az_params* p, int device, b2az_engine** out);
int b2az_destroy(b2az_engine* e);
/* Change the games_to_play budget of a running engine (multi-GPU runs keep ONE global budget, play_manager.cc:506-513:
 * every rank lowers its own target once the all-reduced number of started games reaches the global one). Slots retire
 * when their next game would exceed the budget, exactly as with the value given at creation. */
int b2az_set_games_to_play(b2az_engine* e, uint32_t games_to_play);

/* One pass of PlayManager::play()'s loop body (play_manager.cc:272-599) over EVERY active slot:
 * process_result for the evaluation submitted since the last step, play a move when the search
 * budget is reached (move choice, history capture, re-root, game end / restart), then find_leaf.
 * With B2AZ_EVAL_NN `n_steps` must be 1 and every leaf of the previous step must have been
 * answered (b2az_submit_eval*); leaves go to the leaf batch. With B2AZ_EVAL_RANDOM the evaluator
 * runs inline and `n_steps` passes are fused into one kernel launch. Returns when the work is
 * enqueued on `stream`. */
int b2az_step(b2az_engine* e, uint32_t n_steps, void* stream);

/* build_batch (py_wrapper.cc:449-504), zero-copy flavour: device pointers to the dense
 * canonical batch float32[count][4][6][7] and the slot ids uint32[count] of the leaves produced
 * by the last step. Synchronises `stream` to read `count`. Pointers stay valid until the next
 * b2az_step. */
int b2az_leaf_batch(b2az_engine* e, void* stream, uint32_t* count, const float** canon_dev,
                    const uint32_t** ids_dev);
/* build_batch, legacy host-buffer flavour: copies up to `max` not-yet-taken leaves into host
 * buffers (canonical rows + slot ids), FIFO. */
int b2az_leaf_batch_host(b2az_engine* e, void* stream, uint32_t max, float* canon_host, uint32_t* ids_host,
                         uint32_t* count);

/* The device position cache (the engine's replacement for S3FIFOCache / ShardedS3FIFOCache, s3fifo_cache.h:41-110) key by
 * key, in order: insert = S3FIFOCache::insert (an existing key is never overwritten), find = S3FIFOCache::find (counts a
 * hit or a miss, bumps the entry's frequency; found[i] = 1 and v / pi rows filled on a hit). Keys are opaque non-zero
 * 64-bit position keys; v is float32[n][3], pi float32[n][7]. For tools that share evaluations with an engine and for
 * the tests that hold the table against the reference container. */
int b2az_cache_insert_host(b2az_engine* e, void* stream, const uint64_t* keys_host, const float* v_host,
                           const float* pi_host, uint32_t n);
int b2az_cache_find_host(b2az_engine* e, void* stream, const uint64_t* keys_host, uint32_t n, uint8_t* found_host,
                         float* v_host, float* pi_host);

/* The SEARCHING seat of the first `count` rows of the current leaf batch (the slot's side to move; the leaf's own side to
 * move is in its canonical planes). A PlayManager with several model groups routes row r to the network of group
 * model_groups[seat[r]] (play_manager.cc:577, 598: awaiting_inference_[game.seat_perm[cp]]). */
int b2az_leaf_seats_host(b2az_engine* e, void* stream, uint8_t* seats_host, uint32_t count);
/* The MODEL GROUP of the same rows: seat_perms[slot's permutation][searching seat] (play_manager.cc:577). */
int b2az_leaf_groups_host(b2az_engine* e, void* stream, uint8_t* groups_host, uint32_t count);
/* perm_scores(p) / perm_games_completed(p) for p < n_seat_perms (one entry without permutations): out8[p]. */
int b2az_perm_scores(b2az_engine* e, void* stream, b2az_perm_stats* out8, uint32_t* n_perms_out);

/* update_inferences (play_manager.cc:619-642), zero-copy flavour: v_dev float32[count][3],
 * pi_dev float32[count][7] in leaf-batch row order, count == the leaf count. The buffers are
 * read by the next b2az_step (keep them alive until it has run). */
int b2az_submit_eval(b2az_engine* e, const float* v_dev, const float* pi_dev, uint32_t count);
/* The same pair with NO host synchronisation at all (the whole generation stays stream-ordered and can be
 * captured in a CUDA graph): b2az_leaf_batch_device returns device pointers to the canonical batch
 * float32[concurrent_games][4][6][7] (rows >= *count_dev are unspecified), the slot ids and the row COUNT
 * itself (uint32 in device memory); b2az_submit_eval_all declares that rows [0, count) of v_dev
 * float32[>= concurrent_games][3] / pi_dev float32[>= concurrent_games][7] answer the whole batch. The evaluator
 * simply runs on all concurrent_games rows. */
int b2az_leaf_batch_device(b2az_engine* e, void* stream, const float** canon_dev, const uint32_t** ids_dev,
                           const uint32_t** count_dev);
/* The legal-move masks of the same leaf batch, uint8[B][7] in device memory (row i belongs to row i of the canonical
 * batch): the valid-move half of the zero-copy feed (SURVEY.md 8b "additive exports"). Enqueued on `stream`. */
int b2az_leaf_valid_device(b2az_engine* e, void* stream, const uint8_t** valid_dev);
int b2az_submit_eval_all(b2az_engine* e, const float* v_dev, const float* pi_dev);
/* update_inferences, legacy flavour: row i answers slot ids_host[i]; may be called several times
 * with disjoint subsets. */
int b2az_submit_eval_host(b2az_engine* e, void* stream, const uint32_t* ids_host, const float* v_host,
                          const float* pi_host, uint32_t count);

/* build_history_batch (py_wrapper.cc:393-424): pops up to `max` finished training samples in the
 * order the reference's history_ queue would hold them; canonical float32[n][4][6][7],
 * v float32[n][3], pi float32[n][7]. dst_is_device selects device or host destination. */
int b2az_drain_history(b2az_engine* e, void* stream, uint32_t max, float* canon, float* v, float* pi,
                       int dst_is_device, uint32_t* count);

/* build_history_batch followed by game_runner.exploit_symmetries (game_runner.py:1050-1144), on the device: every
 * popped sample is written together with its symmetric images in GameState::symmetries order (Connect4: the
 * sample, then its mirror image — connect4_gs.cc:151-170), i.e. 2 * *count rows; `max` counts SAMPLES, the
 * buffers must hold 2 * max rows. *count = samples popped. */
int b2az_drain_history_sym(b2az_engine* e, void* stream, uint32_t max, float* canon, float* v, float* pi,
                           int dst_is_device, uint32_t* count);

/* The same drain OVERLAPPED with the next step (build_history_batch runs on its own thread next to play() in the
 * reference, game_runner.py:729-745). b2az_history_mark: on the stream the steps run on, after step k has been enqueued —
 * an asynchronous snapshot of the sample counters, returns at once. b2az_drain_history_marked: on a SECOND stream — waits
 * for the mark only, expands and copies the samples up to it while step k + 1 (already enqueued on the first stream)
 * runs, and returns when the caller's buffers are filled. Do not mix with b2az_drain_history within one run. */
int b2az_history_mark(b2az_engine* e, void* stream);
int b2az_drain_history_marked(b2az_engine* e, void* stream2, uint32_t max, float* canon, float* v, float* pi,
                              int dst_is_device, uint32_t* count);

int b2az_get_stats(b2az_engine* e, void* stream, b2az_stats* out);

/* GameData / MCTS peeks for parity tests (py_wrapper.cc:265-288 game_data(i); mcts.h:101-115
 * counts/root_q_values/root_value/depth/root_n). state89 = Connect4GS::to_bytes layout
 * (connect4_gs.cc:172-178). Any pointer may be NULL. */
int b2az_peek(b2az_engine* e, void* stream, uint32_t game, uint32_t seat, uint8_t* state89, uint32_t* counts7,
              float* root_q7, float* root_value3, uint32_t* depth, uint32_t* root_n, float* root_policy7);

/* Bitboard game kernels on a batch of positions (GameState::play_move / valid_moves / scores /
 * canonicalized, connect4_gs.cc:39-149). boards_host: int8[n][2][6][7]; players/turns per
 * position; moves: one move per position or 0xFFFFFFFF for "no move". Outputs (host, any may be
 * NULL): boards_out int8[n][84], players_out, valid uint8[n][7], scores float[n][3] with
 * terminal[n] = 0/1, canonical float[n][168], status[n] = 0 or B2AZ_EMOVE. */
int b2az_c4_batch(int device, uint32_t n, const int8_t* boards_host, const uint8_t* players, const uint32_t* turns,
                  const uint32_t* moves, int8_t* boards_out, uint8_t* players_out, uint8_t* valid, float* scores,
                  uint8_t* terminal, float* canonical, int32_t* status);

#define B2AZ_TAFL_BRANDUBH 0   /* 7x7,   686 actions, 7 canonical planes (brandubh_gs.h:27-62) */
#define B2AZ_TAFL_OPENTAFL 1   /* 11x11, 2662 actions, 8 planes (opentafl_gs.h:18-50) */
#define B2AZ_TAFL_TAWLBWRDD 2  /* 11x11, 2662 actions, 7 planes (tawlbwrdd_gs.h:20-51) */

/* Tafl game kernels on a batch of game transcripts: {Brandubh,OpenTafl,Tawlbwrdd}GS::play_move / valid_moves /
 * scores / canonicalized / the repetition table (brandubh_gs.cc:225-289, 338-427, 441-537; opentafl_gs.cc:
 * 154-276, 295-428, 441-582; tawlbwrdd_gs.cc:141-213, 221-330, 340-440) replayed on the device from the start
 * position, one warp per game. S = board side, A = 2*S^3 actions, P = canonical planes. moves uint16[n][max_len]
 * (move id = (h*S+w)*2S + (row slide ? new_w : S+new_h), tafl_helper.h:7-14), lens[n]. Outputs (host, any may be
 * NULL) hold the position after k = 0..lens[i] moves at row i*(max_len+1)+k: boards int8[..][3][S][S] (king /
 * defenders / attackers), players, turns, reps (current_repetition_count_), terminal (0 = scores() is nullopt,
 * else 1 + index of the winner, 3 = draw), n_valid (number of legal moves), valid uint8[..][A],
 * canonical float[..][P][S][S]; status[n] = 0 or B2AZ_EMOVE at the first move the reference would have thrown
 * on (rows from there on are unspecified). */
int b2az_tafl_replay(int device, uint32_t game, uint32_t n, uint32_t max_len, uint32_t max_turns,
                     const uint16_t* moves, const uint32_t* lens, int8_t* boards, uint8_t* players, uint32_t* turns,
                     uint8_t* reps, uint8_t* terminal, uint32_t* n_valid, uint8_t* valid, float* canonical,
                     int32_t* status);

/* b2az_tafl_replay with every buffer already in device memory (zero-copy flavour: what a device-resident
 * self-play loop and the throughput measurement use): hist_dev is scratch of n*(max_len+2)*48 bytes for the
 * repetition histories; output pointers may be NULL; the launch is enqueued on `stream`. */
int b2az_tafl_replay_device(uint32_t game, uint32_t n, uint32_t max_len, uint32_t max_turns, const uint16_t* moves_dev,
                            const uint32_t* lens_dev, void* hist_dev, int8_t* boards_dev, uint8_t* terminal_dev,
                            uint32_t* n_valid_dev, uint8_t* valid_dev, float* canonical_dev, int32_t* status_dev,
                            void* stream);

/* The same game kernels on a batch of ARBITRARY positions (the 7-argument GameState constructors, e.g.
 * brandubh_gs.h:124-151): boards int8[n][3][S][S], players, turns, reps (current_repetition_count_) per position.
 * Outputs for the position itself: terminal, n_valid, valid uint8[n][A], canonical float[n][P][S][S]; and, when
 * moves != NULL and moves[i] != 0xFFFFFFFF, play_move(moves[i]) WITHOUT the repetition bookkeeping: boards_out
 * int8[n][3][S][S], captured_any[n] (a capture clears the reference's repetition table), status[n] = 0 or
 * B2AZ_EMOVE. One warp per position. */
int b2az_tafl_positions(int device, uint32_t game, uint32_t n, uint32_t max_turns, const int8_t* boards,
                        const uint8_t* players, const uint32_t* turns, const uint8_t* reps, const uint32_t* moves,
                        uint8_t* terminal, uint32_t* n_valid, uint8_t* valid, float* canonical, int8_t* boards_out,
                        uint8_t* captured_any, int32_t* status);

/* GameState::symmetries for the tafl games (tafl_helper::eightSym, tafl_helper.h:16-149): for each of n samples
 * (canonical float32[n][P][S][S], v float32[n][3], pi float32[n][A]; host pointers) the eight symmetric images in the
 * reference's order [base, rot90, rot180, rot270, mirror(base), mirror(rot90), mirror(rot180), mirror(rot270)] —
 * canon_out float32[n][8][P][S][S], v_out float32[n][8][3], pi_out float32[n][8][A]: the augmentation
 * game_runner.exploit_symmetries (game_runner.py:1050-1144) does sample by sample on the host. */
int b2az_tafl_symmetries(int device, uint32_t game, uint32_t n, const float* canon_host, const float* v_host,
                         const float* pi_host, float* canon_out, float* v_out, float* pi_out);

/* ---- Star Gambit (star_gambit_gs.h / .cc). Game ids: B2AZ_SG_GAME(variant) = the variant's own class
 * (StarGambit{Skirmish,Showdown,Clash,Battle}GS: 11x11 or 13x13 grid, 32 planes, 1229 or 1709 actions),
 * B2AZ_SG_UNIFIED(variant) = StarGambitUnifiedGS pinned to the variant (13x13 canvas, 36 planes, 1709 actions). */
#define B2AZ_SG_SKIRMISH 0
#define B2AZ_SG_SHOWDOWN 1
#define B2AZ_SG_CLASH 2
#define B2AZ_SG_BATTLE 3
#define B2AZ_SG_GAME(variant) (10u + (variant))
#define B2AZ_SG_UNIFIED(variant) (20u + (variant))
#define B2AZ_FOREST_CONNECT4 30u /* Connect4 under the wide-tree search API (b2az_forest_*): the `MCTS` class over Connect4GS */
#define B2AZ_SG_UNIFIED_MIX 24u  /* StarGambitUnifiedGS(-1, probs): every new game draws its variant (self-play engine) */
#define B2AZ_SG_STATE_BYTES 200  /* 20 units x 9 B (star_gambit_gs.h:359-371 order), n_units, reserves[2][4], player,
                                    has_taken_action, game_over, winner, variant, 2 pad, turn u32 */

/* Star Gambit game kernels on a batch of transcripts: StarGambitGS::play_move / valid_moves / scores / canonicalized
 * and the position-key history (star_gambit_gs.cc:784-923, 1093-1290, 1313-1382, 1384-1669; Unified 2522-2616)
 * replayed on the device from the start position, one warp per game. moves uint16[n][max_len], lens[n]. Outputs
 * (host, any may be NULL) hold the position after k = 0..lens[i] moves at row i*(max_len+1)+k: states
 * uint8[..][B2AZ_SG_STATE_BYTES], terminal (0 = scores() is nullopt, 1 + winner, 3 = draw), n_valid, valid
 * uint8[..][A], canonical float[..][P][D][D]; status[n] = 0, B2AZ_EMOVE for an id outside the action space. */
int b2az_sg_replay(int device, uint32_t game, uint32_t n, uint32_t max_len, const uint16_t* moves, const uint32_t* lens,
                   uint8_t* states, uint8_t* terminal, uint32_t* n_valid, uint8_t* valid, float* canonical,
                   int32_t* status);
/* The same with every buffer already in device memory (what a device-resident loop and the throughput measurement
 * use): hist_dev is scratch of n*hist_cap*8 bytes for the key histories (hist_cap >= max_len + 2 never overflows);
 * output pointers may be NULL; the launch is enqueued on `stream`. */
int b2az_sg_replay_device(uint32_t game, uint32_t n, uint32_t max_len, const uint16_t* moves_dev,
                          const uint32_t* lens_dev, void* hist_dev, uint32_t hist_cap, uint8_t* states_dev,
                          uint8_t* terminal_dev, uint32_t* n_valid_dev, uint8_t* valid_dev, float* canonical_dev,
                          int32_t* status_dev, void* stream);

/* GameState::symmetries for Star Gambit (identity + the NW-axis mirror, star_gambit_gs.cc:1671-1805; Unified 2623-2727) on a
 * batch of n samples: canon float32[n][P][D][D], v float32[n][3], pi float32[n][A] -> canon_out [n][2][P][D][D], v_out
 * [n][2][3], pi_out [n][2][A] (every sample followed by its mirror image: the order game_runner.exploit_symmetries writes,
 * game_runner.py:1050-1144). device_pointers != 0: every buffer is device memory and the launch is enqueued on `stream`
 * (e.g. straight from the self-play engine's sample ring); fp16_out != 0: the outputs are IEEE half (the form
 * game_runner.save_compressed stores, game_runner.py:200-210). */
int b2az_sg_symmetries(int device, uint32_t game, uint32_t n, const float* canon, const float* v, const float* pi,
                       void* canon_out, void* v_out, void* pi_out, int device_pointers, int fp16_out, void* stream);

/* ---- Batched single-tree MCTS over the tafl games ("forest"): the reference's `MCTS` class (mcts.h:50-150, bound at
 * py_wrapper.cc:192-220) for n_trees trees at once, one warp per tree on the device. Tree i starts at the game's
 * start position and draws from pcg32(seed + i) — a reference MCTS driven after MCTS::seed_thread_rng(seed + i).
 * Mirrors MCTS(cpuct, num_players = 2, num_moves, epsilon, root_policy_temp, fpu_reduction, relative_values,
 * root_fpu_zero, shaped_dirichlet, gumbel_enabled, gumbel_m, gumbel_c_visit, gumbel_c_scale, gumbel_full): this
 * version implements PUCT and Gumbel root search, root policy temperature and (shaped) Dirichlet noise, with
 * relative_values == 0 and gumbel_full == 0, and rejects anything else. */
typedef struct b2az_forest_params {
  uint32_t game;                 /* B2AZ_TAFL_* */
  uint32_t n_trees;
  uint32_t max_turns;            /* the game's max_turns (BrandubhGS(max_turns) ...) */
  uint32_t words_per_tree;       /* node slab per tree in 32-bit words (1 + 7k words per expanded node); 0 = 2^20 */
  float cpuct, fpu_reduction, epsilon, root_policy_temp;
  uint8_t root_fpu_zero, relative_values, gumbel_enabled, gumbel_full;
  uint32_t gumbel_m;             /* PlayParams::gumbel_m (16) */
  uint64_t seed;
  float gumbel_c_visit, gumbel_c_scale;  /* 50, 1 */
  uint8_t shaped_dirichlet;
  uint8_t debug_serial_shuffle;  /* diagnostics: std::shuffle draws one after the other on one lane (same results) */
  uint8_t pad_[2];
  uint32_t max_in_flight;        /* WU-UCT: most pending leaves per tree (find_leaf_batched); 0 = batched calls off */
} b2az_forest_params;
typedef struct b2az_forest b2az_forest;
int b2az_forest_create(const b2az_forest_params* p, int device, b2az_forest** out);
int b2az_forest_destroy(b2az_forest* f);
/* MCTS::find_leaf(gs) (mcts.cc:462-498) for every tree: PUCT descent from the tree's root position, expansion of
 * the new node (terminal test, legal moves, std::shuffle). *canon_dev = DEVICE pointer to the leaves' canonical
 * planes float32[n_trees][P][S][S] (the evaluator's input); b2az_forest_leaf_canon_host copies them out (all
 * max(1, max_in_flight) slots: float32[slots][n_trees][P][S][S]). */
int b2az_forest_find_leaf(b2az_forest* f, void* stream, const float** canon_dev);
int b2az_forest_leaf_canon_host(b2az_forest* f, void* stream, float* canon_host);
/* MCTS::process_result(gs, value, pi, root_noise_enabled) (mcts.cc:500-555): v float32[n_trees][3],
 * pi float32[n_trees][A], device or host pointers. */
int b2az_forest_process_result(b2az_forest* f, void* stream, const float* v_dev, const float* pi_dev,
                               int root_noise_enabled);
int b2az_forest_process_result_host(b2az_forest* f, void* stream, const float* v_host, const float* pi_host,
                                    int root_noise_enabled);
/* n_sims x (find_leaf -> dumb_eval (game_state.h:160-173) -> process_result) fused in one launch. */
int b2az_forest_simulate(b2az_forest* f, void* stream, uint32_t n_sims, int root_noise_enabled);
/* MCTS::apply_root_policy_temp() then (add_noise != 0 and epsilon > 0) MCTS::add_root_noise() on every tree whose
 * root has been visited: what PlayManager does to the reused root after a move (play_manager.cc:546-553). */
int b2az_forest_root_noise(b2az_forest* f, void* stream, int add_noise);
/* WU-UCT (mcts.cc:752-851): find_leaf_batched appends one pending leaf per tree (virtual loss: ++n_in_flight along
 * the path; at most max_in_flight of them); *canon_dev = DEVICE pointer to float32[max_in_flight][n_trees][P][S][S],
 * slot i = the i-th call since the last reset. process_result_batched(leaf_index, ...) answers slot leaf_index of
 * every tree (v float32[n_trees][3], pi float32[n_trees][A]; host_pointers selects host or device buffers);
 * reset_batch forgets the list (in_flight_.clear()). simulate_batched fuses n_rounds x (width find_leaf_batched,
 * width process_result_batched with dumb_eval, reset_batch) into one launch. */
int b2az_forest_find_leaf_batched(b2az_forest* f, void* stream, const float** canon_dev);
int b2az_forest_process_result_batched(b2az_forest* f, void* stream, uint32_t leaf_index, const float* v, const float* pi,
                                       int root_noise_enabled, int host_pointers);
int b2az_forest_simulate_batched(b2az_forest* f, void* stream, uint32_t n_rounds, uint32_t width);
int b2az_forest_reset_batch(b2az_forest* f, void* stream);
/* MCTS::set_gumbel_num_sims(n) (mcts.cc:175-178) on every tree — call before each move's search, like
 * PlayManager does (play_manager.cc:531-539); n == 0 = PUCT for that search. */
int b2az_forest_set_gumbel_num_sims(b2az_forest* f, void* stream, uint32_t n);
/* MCTS::gumbel_final_action() (uint32[n_trees]; 0xFFFFFFFF where the search never initialised: the reference then
 * falls back to pick_move(probs(0))) and MCTS::gumbel_improved_policy() (float32[n_trees][A]) (mcts.cc:336-401).
 * Host pointers, either may be NULL. */
int b2az_forest_gumbel_result(b2az_forest* f, void* stream, uint32_t* action_host, float* policy_host);
/* MCTS::probs(temp) (mcts.cc:575-618: visit-count policy with temperature; temp == 0 = uniform over the most
 * visited moves; the tempered priors when nothing has a visit) into probs_host float32[n_trees][A] (may be NULL),
 * and with pick_move != 0 MCTS::pick_move(probs) (mcts.cc:717-735, exactly one draw from the tree's generator) into
 * moves_host uint32[n_trees] — PlayManager's acting rule (play_manager.cc:372-381). pruned != 0: MCTS::probs_pruned
 * (mcts.cc:620-674, the policy target under policy_target_pruning, play_manager.cc:418-421). */
int b2az_forest_probs(b2az_forest* f, void* stream, float temp, int pruned, int pick_move, float* probs_host,
                      uint32_t* moves_host);
/* Greedy self-play step on the device: every tree whose root is expanded and not terminal plays its most visited
 * move (argmax of MCTS::counts(), lowest move id on ties) through update_root + play_move. No host round trip. */
int b2az_forest_advance(b2az_forest* f, void* stream);
/* MCTS::update_root(gs, move) (mcts.cc:151-173) followed by gs.play_move(move) on the tree's root position;
 * moves_host[n_trees], 0xFFFFFFFF = leave that tree alone. */
int b2az_forest_update_root(b2az_forest* f, void* stream, const uint32_t* moves_host);
/* MCTS::counts / root_q_values (mcts.cc:557-573): uint32[n_trees][A], float32[n_trees][A]; info uint32[n_trees][16] =
 * depth, root n, root children, root terminal code, root player, turn, repetition count, error bits, slab words
 * used, root v (bits), total_leaf_depth, side to move, MCTS::root_value() win / loss / draw (float bits, mcts.h:78-100),
 * in_flight_count. Any pointer may be NULL. */
int b2az_forest_counts(b2az_forest* f, void* stream, uint32_t* counts_host, float* q_host, uint32_t* info_host);
/* MCTS::principal_variation(depth) (mcts.cc:676-715): moves_host uint32[n_trees][depth], len_host[n_trees]. */
int b2az_forest_principal_variation(b2az_forest* f, void* stream, uint32_t depth, uint32_t* moves_host, uint32_t* len_host);
/* The moves from the root to a pending leaf (MCTS::path_, or in-flight leaf `slot` of a WU-UCT forest; slot < 0 = the
 * plain find_leaf's leaf): moves_host uint32[n_trees][96], len_host[n_trees]. A caller that holds the root GameState
 * replays them to obtain the leaf position find_leaf returns (py_wrapper.cc:199). */
int b2az_forest_leaf_path(b2az_forest* f, void* stream, int slot, uint32_t* moves_host, uint32_t* len_host);
/* MCTS::apply_root_policy_temp (mcts.cc:448-460) and / or MCTS::add_root_noise (mcts.cc:403-446) on every expanded root. */
int b2az_forest_root_ops(b2az_forest* f, void* stream, int apply_temp, int add_noise);
/* Start tree `tree` from another position than the game's initial one (the GameState a caller hands to
 * MCTS::find_leaf): `state` is the engine's position record (TaflState 72 B / B2AZ_SG_STATE_BYTES) as the host GameState
 * classes of the alphazero module hold it, `hist` its repetition keys. Only before the tree's first search. */
int b2az_forest_set_root(b2az_forest* f, uint32_t tree, const void* state, uint32_t state_bytes, const void* hist,
                         uint32_t hist_count);
/* The inverse: the root position of `tree` (GameData::gs() of a PlayManager slot, py_wrapper.cc:265-288): the position
 * record into state[state_bytes], up to hist_cap repetition keys into hist, their number into *hist_count. Synchronises. */
int b2az_forest_get_root(b2az_forest* f, uint32_t tree, void* state, uint32_t state_bytes, void* hist, uint32_t hist_cap,
                         uint32_t* hist_count);

/* ---- PlayManager::play (play_manager.cc:258-600) over the tafl games on the device: n_games game slots, each with the
 * two seats' search trees (GameData::mcts[0..1]) and ONE pcg32 stream, playing games_per_slot games one after the
 * other. Slot g reproduces the unmodified reference PlayManager with concurrent_games = 1, games_to_play =
 * games_per_slot, mcts_visits = {visits, visits}, self_play, no playout cap, no resignation, run on one thread after
 * MCTS::seed_thread_rng(forest.seed + g): moves, training samples (canonical, outcome, policy target: the Gumbel
 * improved policy / probs_pruned(1) / probs(1), play_manager.cc:417-435), scores and metrics, bit for bit.
 * forest.n_trees is ignored (2 * n_games trees are made); forest.max_in_flight must be 0. */
typedef struct b2az_tafl_selfplay_params {
  b2az_forest_params forest;     /* game (B2AZ_TAFL_*, B2AZ_SG_GAME / B2AZ_SG_UNIFIED), max_turns (Star Gambit: how many
                                    training samples of one game are staged, e.g. 768; a longer game keeps its last ones), search parameters (cpuct, epsilon,
                                    Gumbel, relative_values ...), seed, slab size */
  uint32_t n_games;              /* PlayParams::concurrent_games */
  uint32_t games_per_slot;       /* games every slot plays before it retires (games_to_play = n_games * games_per_slot) */
  uint32_t visits;               /* PlayParams::mcts_visits (both seats) */
  float start_temp, final_temp, temp_decay_half_life;  /* play_manager.cc:285-302 */
  uint8_t history_enabled, policy_target_pruning, tree_reuse, pad_;
  uint32_t hist_capacity;        /* rows of the training-sample ring between drains; 0 = n_games * max_turns */
  uint32_t seat_visits[2];       /* per-seat search budget (seat_visits, play_manager.cc:70-80); 0 = visits */
  uint32_t seat_cap_visits[2];   /* per-seat fast-search budget (seat_cap_visits, :82-90); 0 = playout_cap_depth */
  uint32_t playout_cap_depth;    /* PlayParams::playout_cap_depth (25) */
  float playout_cap_percent;     /* PlayParams::playout_cap_percent */
  float resign_percent, resign_playthrough_percent;  /* play_manager.cc:305-333 */
  uint8_t playout_cap_randomization; /* fast searches: no sample, no root noise, PUCT acting (play_manager.cc:523-553) */
  uint8_t fast_search_uses_gumbel;
  uint8_t pad2_[2];
  uint32_t n_variant_half_life;  /* PlayParams::temp_decay_half_life_by_variant (play_manager.cc:289-296): entries used */
  float variant_half_life[4];    /* indexed by GameState::get_variant_id() (StarGambitUnifiedGS: the variant) */
  float variant_probs[4];        /* game B2AZ_SG_UNIFIED_MIX: StarGambitUnifiedGS's variant weights (all 0 = 0.25 each) */
  uint32_t cache_entries;        /* PlayParams::max_cache_size: entries of the device position cache (EvalType::NN form:
                                    a slot keeps simulating while its leaves hit, play_manager.cc:589-594); 0 = no cache */
  /* Seat permutations and model groups (PlayParams::seat_perms / model_groups, play_manager.cc:24-90, 213-221; what
   * game_runner.play_past and tournament.py build): seat_perms[p][seat] = the model group (0 or 1) that searches for
   * `seat` in the games of permutation p. Slot g plays permutation g % n_seat_perms in every one of its games (the
   * reference's round-robin hand-out when the slots finish in order), so n_games must be a multiple of n_seat_perms.
   * 0 = one seating, one model group. One position-cache table serves both groups (the group is part of the key). */
  uint32_t n_seat_perms;         /* <= 8 */
  uint8_t seat_perms[8][2];
  uint32_t perm_seat_visits[8][2];      /* seat_visits_[p][seat] (play_manager.cc:70-80); 0 = seat_visits[seat] */
  uint32_t perm_seat_cap_visits[8][2];  /* seat_cap_visits_[p][seat] (:82-90); 0 = seat_cap_visits[seat] */
  uint8_t group_random[2];       /* model group i is EvalType::RANDOM next to an NN group (eval_types_[group],
                                    play_manager.cc:578-587): its searches run dumb_eval on the device and never
                                    show up in the leaf batch */
  uint8_t has_seat_search;       /* 1: the per-(permutation, seat) search settings below replace forest.epsilon /
                                    root_policy_temp / root_fpu_zero / gumbel_* — what make_mcts hands every seat's MCTS
                                    (seat_epsilon_ ... seat_gumbel_full_, play_manager.cc:92-164, 602-617) — and the
                                    per-seat resign rule is on (seat_resign_threshold_ / _consecutive_, :335-366) */
  uint8_t pad4_;
  float seat_epsilon[8][2], seat_root_temp[8][2];
  uint8_t seat_root_fpu_zero[8][2], seat_gumbel_enabled[8][2], seat_gumbel_full[8][2];
  uint32_t seat_gumbel_m[8][2];
  float seat_gumbel_c_visit[8][2], seat_gumbel_c_scale[8][2];
  float seat_resign_threshold[8][2];     /* -2 = off */
  uint32_t seat_resign_consecutive[8][2];
} b2az_tafl_selfplay_params;
typedef struct b2az_tafl_selfplay_slot {  /* per-slot share of PlayManager's counters (play_manager.cc:462-505) */
  uint32_t active, games_started, games_completed, pending;
  uint32_t move_count, full_move_count;                          /* current game */
  uint32_t total_move_count, total_full_move_count, game_length; /* finished games */
  uint32_t picked, error, capped;                                /* error: 1 = sample ring full (samples dropped) */
  double g_leaf_depth, g_entropy, g_valid_moves;                 /* current game */
  double leaf_depth, entropy, valid_moves;                       /* total_avg_leaf_depth_, total_search_entropy_, total_valid_moves_ */
  unsigned long long simulations;
  float scores[3];                                               /* scores_: seat 0 wins, seat 1 wins, draws */
  uint32_t playthrough;
  uint32_t fast_move_count, total_fast_move_count;               /* capped (fast) searches */
  double g_fast_leaf_depth, g_fast_entropy, fast_leaf_depth, fast_entropy;
  float resign_scores[3];                                        /* resign_scores_ */
  uint16_t resign_streak[2];                                     /* GameData::resign_streak (per-seat resign rule) */
  uint64_t coin_state, coin_inc;                                 /* the slot's coin stream (playout cap, resign playthrough) */
} b2az_tafl_selfplay_slot;
typedef struct b2az_tafl_selfplay b2az_tafl_selfplay;
int b2az_tafl_selfplay_create(const b2az_tafl_selfplay_params* p, int device, b2az_tafl_selfplay** out);
int b2az_tafl_selfplay_destroy(b2az_tafl_selfplay* sp);
/* EvalType::RANDOM: n_moves x (visits simulations with dumb_eval fused, then the move step) for every active slot,
 * two launches per move, no host synchronisation unless active_out (host, number of slots still cycling) is given. */
int b2az_tafl_selfplay_play(b2az_tafl_selfplay* sp, void* stream, uint32_t n_moves, uint32_t* active_out);
/* EvalType::NN, one simulation per call pair: find_leaf writes the active slots' leaf positions (row g = slot g of
 * DEVICE float32[n_games][P][S][S], the evaluator's input, zero copy); process_result takes (v float32[n_games][3],
 * pi float32[n_games][A]) in the same row order (device pointers, or host ones with host_pointers != 0) and plays the
 * move of every slot whose search has reached `visits` simulations. */
int b2az_tafl_selfplay_find_leaf(b2az_tafl_selfplay* sp, void* stream, const float** canon_dev);
int b2az_tafl_selfplay_process_result(b2az_tafl_selfplay* sp, void* stream, const float* v, const float* pi, int host_pointers,
                                      uint32_t* active_out);
/* PlayManager::build_history_batch (py_wrapper.cc:393-424): up to max_rows of the waiting samples, oldest first (each
 * game's moves last first, like history_), with the slot they came from; the rest stay for the next call. */
int b2az_tafl_selfplay_drain_history(b2az_tafl_selfplay* sp, void* stream, uint32_t max_rows, float* canon_host, float* v_host,
                                     float* pi_host, uint32_t* slot_host, uint32_t* n_out);
/* PlayManager's per-variant tables (variant_scores_, variant_metrics_: play_manager.cc:468-484, play_manager.h:218-275) for
 * games with variants (StarGambitUnifiedGS: B2AZ_SG_UNIFIED(v), B2AZ_SG_UNIFIED_MIX): out4[v], sums over completed games. */
typedef struct b2az_variant_stats {
  float scores[3];
  uint32_t games_completed;
  uint32_t game_length, total_move_count, full_move_count, fast_move_count;
  double leaf_depth, entropy, valid_moves, fast_leaf_depth, fast_entropy;
} b2az_variant_stats;
int b2az_tafl_selfplay_variant_stats(b2az_tafl_selfplay* sp, void* stream, b2az_variant_stats* out4);
/* b2az_get_stats for this engine: PlayManager's counters and getters (play_manager.h:173-180, 288-316) over all slots. */
int b2az_tafl_selfplay_get_stats(b2az_tafl_selfplay* sp, void* stream, b2az_stats* out);
/* The reference-API flavour of one simulation with HOST buffers (build_batch / update_inferences, py_wrapper.cc:449-504,
 * play_manager.cc:619-642): find_leaf for every active slot, their canonical planes compacted into canon_host
 * float32[n][P][S][S] with the slot ids in ids_host[n]; then the answers v float32[n][3], pi float32[n][A] (row i for
 * slot ids[i]; every active slot must be answered): process_result + the move of every slot whose search is complete. */
int b2az_tafl_selfplay_leaf_batch_host(b2az_tafl_selfplay* sp, void* stream, uint32_t max_rows, float* canon_host,
                                       uint32_t* ids_host, uint32_t* n_out);
int b2az_tafl_selfplay_submit_eval_host(b2az_tafl_selfplay* sp, void* stream, const uint32_t* ids, const float* v, const float* pi,
                                        uint32_t n);
/* The model group of every row of the last b2az_tafl_selfplay_leaf_batch_host (row i = slot ids[i]): the group that
 * searches for the slot's side to move under the slot's seat permutation (awaiting_inference_[game.seat_perm[cp]],
 * play_manager.cc:577, 598). All 0 with one model group. */
int b2az_tafl_selfplay_leaf_groups_host(b2az_tafl_selfplay* sp, uint8_t* groups_host, uint32_t n);
/* PlayManager's per-permutation tables (perm_scores_, variant_perm_scores_: play_manager.cc:205-255, 466-474): out[p] for
 * p < n_seat_perms (one entry when there are no permutations). */
int b2az_tafl_selfplay_perm_stats(b2az_tafl_selfplay* sp, void* stream, b2az_perm_stats* out8, uint32_t* n_perms_out);
/* slots_host[n_games]; tree_errors_host[2 * n_games] = the trees' sticky error bits (see b2az_forest_counts). */
int b2az_tafl_selfplay_slots(b2az_tafl_selfplay* sp, void* stream, b2az_tafl_selfplay_slot* slots_host, uint32_t* tree_errors_host);
/* GameData::gs() of slot `slot` (play_manager.h:33-58): b2az_forest_get_root of the slot's first tree (both seats' trees
 * hold the same position). */
int b2az_tafl_selfplay_root_state(b2az_tafl_selfplay* sp, uint32_t slot, void* state, uint32_t state_bytes, void* hist,
                                  uint32_t hist_cap, uint32_t* hist_count);

#ifdef __cplusplus
}
#endif
#endif /* B2AZ_H_ */
