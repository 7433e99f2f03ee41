/*
 * acmpc_b200.h -- C ABI of the B200-native MPC step (drop-in boundary).
 *
 * The reference (Adelaide-Autonomous-Racing-Kit/ac-mpc) is pure Python and has no FFI layer;
 * the boundary a maintainer would bind is the object API of `acmpc.control`:
 *
 *   acmpc_create            <-> build_mpc / SpatialMPC.__init__      controller.py:19-29, spatial_mpc.py:21-58
 *   acmpc_solve_batch_*     <-> SpatialMPC.get_control               spatial_mpc.py:170-217
 *                               (construct_waypoints :125-154, compute_speed_profile :89-123,
 *                                SpatialBicycleModel.t2s/linearise/s2t dynamics.py:23-103,
 *                                ControlSolver.solve solvers/control.py:15-106,
 *                                SpeedProfileSolver.solve solvers/speed_profile.py:15-86,131-150,
 *                                and the `osqp` solve those two call)
 *   acmpc_outputs fields    <-> the attributes get_control mutates    spatial_mpc.py:193-212
 *   acmpc_destroy           <-> object lifetime
 *
 * Plain pointers and sizes only; no torch / C++ types.  All floating point is IEEE binary64.
 * Per-instance solver failure is reported in `status[]`, never through the return code; the return
 * code is non-zero only for API misuse or CUDA errors (see acmpc_last_error).
 */
#ifndef ACMPC_B200_H
#define ACMPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACMPC_ABI_VERSION 2

/* OSQP status codes written to status[] / status_speed[] ("solved" == 1, spatial_mpc.py:115,193) */
#define ACMPC_SOLVED 1
#define ACMPC_SOLVED_INACCURATE 2
#define ACMPC_PRIMAL_INFEASIBLE_INACCURATE 3
#define ACMPC_DUAL_INFEASIBLE_INACCURATE 4
#define ACMPC_MAX_ITER_REACHED (-2)
#define ACMPC_PRIMAL_INFEASIBLE (-3)
#define ACMPC_DUAL_INFEASIBLE (-4)
#define ACMPC_NON_CVX (-7)
#define ACMPC_UNSOLVED (-10)

/* return codes */
#define ACMPC_OK 0
#define ACMPC_ERR_INVALID 1     /* bad argument / unsupported horizon */
#define ACMPC_ERR_CUDA 2        /* CUDA runtime error, text in acmpc_last_error */
#define ACMPC_ERR_NO_DEVICE 3   /* no sm_100 device: there is NO CPU fallback */

#define ACMPC_MIN_HORIZON 4
#define ACMPC_MAX_HORIZON 128

typedef struct acmpc_config {
    int32_t horizon;            /* H = config["horizon"]; n = H-1 stages          spatial_mpc.py:27 */
    int32_t max_iter;           /* MAX_SOLVER_ITERATIONS = 4000                   spatial_mpc.py:17 */
    /* config["speed_profile_constraints"]                                        spatial_mpc.py:36 */
    double v_min, v_max, a_min, a_max, ay_max, ki_min, end_velocity;
    int32_t has_end_velocity;   /* 0 <=> end_velocity: null                       speed_profile.py:42 */
    int32_t reserved0;
    double step_cost[3];        /* Q  = diag(step_cost)   e_y, e_psi, t           solvers/control.py:126 */
    double r_term[2];           /* R  = diag(r_term)      v, kappa_cmd            solvers/control.py:127 */
    double final_cost[3];       /* QN = diag(final_cost)                          solvers/control.py:128 */
    /* vehicle_data.vehicle_data.wheelbase / .width, max_steering_angle()         dynamics.py:11-13 */
    double wheelbase, width, delta_max;
    /* build-time velocity limits that become the input bounds min_u/max_u        controller.py:20-24 */
    double input_v_min, input_v_max;
    /* OSQP settings; the reference leaves all of them at the library defaults */
    double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
    int32_t scaling;               /* Ruiz passes, 10 */
    int32_t check_termination;     /* 25 */
    int32_t adaptive_rho;          /* 1 */
    int32_t adaptive_rho_interval; /* fixed iteration interval, default 50 */
    /* OSQP >= 1.0 termination semantics (ABI 2).  0 (default) = OSQP 0.6.x: "solved" <=> primal and dual residual
     * tests.  1 = OSQP 1.x `check_dualgap`: additionally |x'Px + q'x + SC(y)| <= eps_abs + eps_rel * max(|x'Px|, |q'x|,
     * |SC(y)|), SC(y) = u'max(y,0) + l'min(y,0) over the finite bounds.  requirements.txt:2 of the reference does not pin
     * the wheel, so a maintainer may be running either; tools/pin_osqp.py reports which setting their wheel matches. */
    int32_t check_dualgap;
    int32_t reserved1;
} acmpc_config;

/* Result arrays, one slice per instance.  Any pointer may be NULL (field not written).
 * n = H-1.  For the *_device entry point these are device pointers, for *_host host pointers. */
typedef struct acmpc_outputs {
    double *controls;       /* [B,2,n] row 0 = v_k, row 1 = delta_k = atan(kappa_cmd_k * L)   projected_control */
    double *prediction;     /* [B,n,2] (X_k, Y_k) world-frame rollout (s2t)                  current_prediction */
    double *cum_time;       /* [B,n]   t_k                                                    cum_time */
    double *states;         /* [B,H,3] spatial states x_0..x_{H-1} = dec.x[:3H] */
    double *v_ref;          /* [B,n]   speed profile (zeros if the speed QP was not "solved") */
    double *cost;           /* [B]     OSQP info.obj_val of the control QP = 1/2 x'Px + q'x */
    double *pri_res;        /* [B]     control QP primal residual (unscaled, inf-norm) */
    double *dua_res;        /* [B]     control QP dual residual */
    int32_t *status;        /* [B]     control QP status */
    int32_t *status_speed;  /* [B]     speed-profile QP status */
    int32_t *iters;         /* [B,2]   ADMM iterations: speed QP, control QP */
    int32_t *rho_updates;   /* [B,2]   refactorisations caused by adaptive rho */
    double *waypoints;      /* [B,7,n] ReferencePath rows xs ys psis kappas distances widths velocities  reference_path */
    double *derived;        /* [B,3,n-1] rows times = diff(t), accelerations = diff(e_y) / times (sic), steer_rates =
                             * diff(e_psi) / times over x_0..x_{n-1}                        spatial_mpc.py:208-211 (ABI 2) */
} acmpc_outputs;

typedef struct acmpc_handle acmpc_handle;

int32_t acmpc_abi_version(void);

/* Fill *cfg with the OSQP defaults, the Monza racing block (configs/monza.yaml:67-81) and the
 * documented synthetic vehicle constants (wheelbase 2.65 m, width 1.99 m, delta_max 0.30 rad). */
void acmpc_default_config(acmpc_config *cfg);

/* Create a solver bound to CUDA device `device`.  Fails with ACMPC_ERR_NO_DEVICE when no CUDA
 * device is present: the product has no CPU path. */
int32_t acmpc_create(const acmpc_config *cfg, int32_t device, acmpc_handle **out);
int32_t acmpc_destroy(acmpc_handle *h);
const char *acmpc_last_error(const acmpc_handle *h);

/* bytes of one instance's warm-start record: the state the reference's three persistent OSQP objects keep
 * between calls (spatial_mpc.py:43-58: speed-profile solver, localised speed-profile solver, control
 * solver) -- scaled x, z, y, the adapted rho and an "object exists" flag each.  A zero-filled record means
 * "no object yet": the first solve of a slot is a cold setup, exactly like the reference's first call. */
int64_t acmpc_warm_stride(const acmpc_handle *h);

/* One MPC step for B independent instances, everything resident on the device.
 *   d_paths   [B,H,3] rows (x, y, width) in the ego frame (x right, y forward)   get_control arg 1
 *   d_offsets [B] lateral offset (get_control arg 3) or NULL (= 0.0)
 *   d_vmax    [B] live speed_profile_constraints["v_max"] (controller.py:241-243) or NULL (= cfg.v_max)
 *   is_localised: get_control arg 2
 *   d_warm    NULL = cold start (x=z=y=0, rho=cfg.rho) ; else [B, acmpc_warm_stride] bytes (8-byte aligned,
 *             zero-filled by the caller before first use), read when warm_valid != 0 and always rewritten
 *             (the reference's persistent OSQP objects: OSQP warm-starts x, y, z and keeps the adapted rho,
 *             while update() re-equilibrates the new data)
 *   stream    cudaStream_t (NULL = default stream).  Asynchronous: no host sync inside (except when the
 *             internal speed-profile hand-over buffer has to grow: first call, or a larger B, with
 *             d_out->v_ref == NULL).  Calls on one handle must be stream-ordered with each other (the
 *             reference object is single-threaded): the work queue of the persistent warps
 *             is per handle.  The device and the host entry points keep separate work queues and order buffers, so they
 *             may be mixed on one handle; they still share the handle's warm-start records only through the pointers
 *             the caller passes. */
int32_t acmpc_solve_batch_device(acmpc_handle *h, int32_t B, const double *d_paths,
                                 const double *d_offsets, const double *d_vmax,
                                 int32_t is_localised, void *d_warm, int32_t warm_valid,
                                 const acmpc_outputs *d_out, void *stream);

/* Multi-GPU completion protocol (ac_mpc_b200/sharded.py, transport "peer").  The output pointers
 * of acmpc_solve_batch_device may be PEER-MAPPED: memory of another GPU of the node, so that the kernels store their
 * results straight into the consumer's buffer over NVLink and no collective moves them afterwards.
 *   acmpc_attach_completion  one-shot, applies to the NEXT acmpc_solve_batch_device call on this handle:
 *       d_flag / flag_value     once every output store of that call is visible system-wide, the control kernel's last
 *                               CTA writes flag_value to *d_flag (a word in the consumer's memory; may be NULL)
 *       d_credit_table[credit_n] device array of addresses of the producers' credit words (<= 32); the call's first kernel
 *                               writes credit_value to each: "the buffers of every step < credit_value are free again"
 *       d_credit_wait / credit_need  producer side: no CTA of the call's first kernel starts before the LOCAL word
 *                               *d_credit_wait is >= credit_need (polled inside the kernel, so consecutive launches
 *                               stay pipelined; NULL = no wait)
 *   acmpc_stream_wait_value32   makes `stream` wait until *d_addr >= value (cuStreamWaitValue32: no SM, no kernel);
 *                               d_addr must be LOCAL device memory of the handle's GPU (the words above are written
 *                               remotely, waited on locally). */
int32_t acmpc_attach_completion(acmpc_handle *h, uint32_t *d_flag, uint32_t flag_value, const uint64_t *d_credit_table,
                                int32_t credit_n, uint32_t credit_value, const uint32_t *d_credit_wait,
                                uint32_t credit_need);
int32_t acmpc_stream_wait_value32(acmpc_handle *h, const uint32_t *d_addr, uint32_t value, void *stream);

/* Same with HOST buffers: copies inputs to the device, runs the kernels, copies every non-NULL
 * output back and synchronises.  keep_warm != 0: instance slot b of consecutive calls with the same B is one
 * persistent solver object (records live on the device inside the handle; a different B or a call with
 * keep_warm == 0 drops them) -- the B = 1 path the drop-in SpatialMPC.get_control uses. */
int32_t acmpc_solve_batch_host(acmpc_handle *h, int32_t B, const double *paths,
                               const double *offsets, const double *vmax, int32_t is_localised,
                               int32_t keep_warm, const acmpc_outputs *out);

/* SpatialMPC.compute_speed_profile(reference_path, is_localised, end_vel) (spatial_mpc.py:89-123) for B ReferencePaths,
 * the speed-profile kernel alone.  waypoints [B,7,n] rows xs ys psis kappas distances widths velocities (paths.py): the
 * kappas and distances rows are read; the velocities row of instance b is WRITTEN ONLY when its QP is "solved" and left
 * untouched otherwise (spatial_mpc.py:115-122).  vmax [B] = the live speed_profile_constraints["v_max"] (NULL = cfg.v_max);
 * has_end_vel / end_vel = the call's end_vel argument (None <=> has_end_vel == 0; ignored when is_localised,
 * speed_profile.py:131-137); the other constraints come from the handle's config.
 * solution [B,n] (may be NULL) = dec.x whatever the status; status [B], iters [B], rho_updates [B] may be NULL.
 * Warm start: the same records as acmpc_solve_batch_* (slot 0 = speed solver, slot 1 = localised speed solver), so a
 * call sequence mixing get_control and compute_speed_profile on one object behaves like the reference's, which shares
 * the two OSQP objects between them (spatial_mpc.py:43-58,101-105).
 * _device: device pointers, asynchronous on `stream`; d_iters / d_rho_updates are [B,2] pairs like acmpc_outputs
 * (column 0 is written).  _host: host pointers, synchronous, iters / rho_updates are [B]. */
int32_t acmpc_speed_profile_batch_device(acmpc_handle *h, int32_t B, double *d_waypoints, const double *d_vmax,
                                         int32_t is_localised, int32_t has_end_vel, double end_vel, void *d_warm,
                                         int32_t warm_valid, double *d_solution, int32_t *d_status, int32_t *d_iters,
                                         int32_t *d_rho_updates, void *stream);
int32_t acmpc_speed_profile_batch_host(acmpc_handle *h, int32_t B, double *waypoints, const double *vmax,
                                       int32_t is_localised, int32_t has_end_vel, double end_vel, int32_t keep_warm,
                                       double *solution, int32_t *status, int32_t *iters, int32_t *rho_updates);

/* SpatialBicycleModel as stand-alone batched calls (inside a step these are fused into the control kernel):
 *   acmpc_t2s_host        dynamics.py:23-40   waypoints[B,3] = (x, y, psi) of the reference waypoint, states[B,3] = (x, y,
 *                                             psi) of the vehicle -> out[B,3] = (e_y, e_psi wrapped to [-pi, pi), t = 0)
 *   acmpc_s2t_host        dynamics.py:42-63   waypoints[B,7,n], states[B,n,3] -> out[B,3,n] rows X, Y, Psi (may be NULL);
 *                                             prediction[B,n,2] (may be NULL) = SpatialMPC.update_prediction =
 *                                             s2t(...)[:-1].T (spatial_mpc.py:156-168)
 *   acmpc_linearise_host  dynamics.py:65-103  waypoints[B,7,n] -> f[B,n,3], A[B,n,3,3], Bm[B,n,3,2] (each may be NULL) */
int32_t acmpc_t2s_host(acmpc_handle *h, int32_t B, const double *waypoints, const double *states, double *out);
int32_t acmpc_s2t_host(acmpc_handle *h, int32_t B, int32_t n, const double *waypoints, const double *states, double *out,
                       double *prediction);
int32_t acmpc_linearise_host(acmpc_handle *h, int32_t B, int32_t n, const double *waypoints, double *f, double *A,
                             double *Bm);

/* Counters of the last call: kernel launches issued (per chunk: speed-profile kernel + control kernel, plus the
 * small ordering kernel for batches of 1024+ with per-instance v_max; the host entry point splits batches of
 * 512+ / 2048+ into 2 / 4 chunks and larger ones into chunks of at most 16384 instances), dynamic shared memory per CTA of the control kernel, threads per CTA,
 * instances per CTA. */
int32_t acmpc_last_launch_info(const acmpc_handle *h, int32_t *n_launches, int32_t *smem_bytes,
                               int32_t *threads_per_cta, int32_t *instances_per_cta);

/* Per-kernel device times.  acmpc_set_profiling(h, 1) makes every following launch record CUDA events on its
 * stream around the two kernels of the step (speed-profile kernel, control kernel); acmpc_collect_kernel_ms
 * synchronises on them and returns the summed durations and the number of launches since the previous
 * collect (at most 256 launches are remembered).  Off by default: it costs three event records per launch. */
int32_t acmpc_set_profiling(acmpc_handle *h, int32_t on);
int32_t acmpc_collect_kernel_ms(acmpc_handle *h, double *speed_ms, double *control_ms, int32_t *launches);

/* FP64 FMA micro-benchmark (dependent-chain-free DFMA loop on every SM): the roofline
 * denominator for this path, MEASURED_PEAKS.json has no FP64 figure.  Returns TFLOP/s. */
int32_t acmpc_fp64_peak_tflops(int32_t device, double *tflops);

/* ---- Whole-track speed profile -------------------------------------------------
 * The start-up path Controller.compute_track_speed_profile (controller.py:49-57): construct_waypoints over the
 * whole centre line (spatial_mpc.py:125-154), then compute_map_speed_profile (spatial_mpc.py:60-87) = ONE
 * speed-profile QP (solvers/speed_profile.py:26-86) with n = M - 1 ~ 10^4 variables and max_iter = 40000, then
 * agent.py:287-302 / :137-143 (savgol_filter(v, 21, 3), [-25, +75) window mean with wrap-around).
 * One cooperative launch: one stage per thread, 512 threads per CTA, so n <= 512 * (SMs of the device). */
#define ACMPC_MAP_MAX_ITER 40000 /* MAX_SOLVER_ITERATIONS_MAP, spatial_mpc.py:16 */

typedef struct acmpc_map_info {
    int32_t status;      /* OSQP status of the QP (ACMPC_SOLVED ...) */
    int32_t iters;       /* ADMM iterations */
    int32_t rho_updates; /* refactorisations caused by adaptive rho */
    int32_t ctas;        /* CTAs of the cooperative launch */
    double pri_res, dua_res, obj_val, rho;
    double kernel_ms;    /* device time of the launch (CUDA events) */
} acmpc_map_info;

/* SpatialMPC.construct_waypoints (spatial_mpc.py:125-154) over an (M,3) host array of (x, y, width):
 * waypoints[7, M-1] = rows xs ys psis kappas distances widths velocities(= 0). */
int32_t acmpc_construct_waypoints_host(acmpc_handle *h, int32_t M, const double *track, double *waypoints);

/* SpatialMPC.compute_map_speed_profile (spatial_mpc.py:60-87) on a ReferencePath: reads the kappas and
 * distances rows of waypoints[7,n] (host), writes the velocities row when the QP is "solved" and leaves it
 * untouched otherwise (spatial_mpc.py:115-123).  v_max = the live speed_profile_constraints["v_max"]; the other
 * constraints come from the handle's config with a_min / ay_max replaced (the deep copy of spatial_mpc.py:80-82).
 * max_iter <= 0 selects ACMPC_MAP_MAX_ITER.  solution (may be NULL) receives dec.x whatever the status. */
int32_t acmpc_map_speed_profile_host(acmpc_handle *h, int32_t n, double *waypoints, double v_max, double ay_max,
                                     double a_min, int32_t max_iter, double *solution, acmpc_map_info *info);

/* Both steps in one launch, straight from the (M,3) track (controller.py:49-57). */
int32_t acmpc_track_speed_profile_host(acmpc_handle *h, int32_t M, const double *track, double v_max,
                                       double ay_max, double a_min, int32_t max_iter, double *waypoints,
                                       double *solution, acmpc_map_info *info);

/* agent.py:300 reference_speeds = savgol_filter(velocities, 21, 3) -> smoothed[n], and agent.py:137-143 for EVERY
 * map index c: window_mean[c] = mean(smoothed[(c - behind .. c + ahead) mod n]) (the reference uses 25 / 75).
 * Either output may be NULL.  n >= 21. */
int32_t acmpc_reference_speeds_host(acmpc_handle *h, int32_t n, const double *velocities, int32_t behind,
                                    int32_t ahead, double *smoothed, double *window_mean);

/* ---- Caller side of the step, batched over B instances ------------------------------
 * ControlProcess._reference_path (controller.py:257-267): perceived centre lines [B,P,2] float32 -> paths [B,H,3]
 * float64 = rows 0::int(P/H) of (x, y) + widths linspace(10, 6, H).  Like the reference (np.stack raises) this is
 * only defined when the stride yields exactly H rows; otherwise ACMPC_ERR_INVALID. */
int32_t acmpc_reference_paths_host(acmpc_handle *h, int32_t B, int32_t P, const float *centrelines, double *paths);

/* ControlProcess._update_shared_memory (controller.py:274-280) into the float32 shared arrays
 * (perception/shared_memory.py:90-104): control_inputs[B,n,2] = projected_control.T, control_cumtime[B,n],
 * predicted_locations[B,n,2].  Any input/output pair may be NULL. */
int32_t acmpc_publish_host(acmpc_handle *h, int32_t B, const double *controls, const double *cum_time,
                           const double *prediction, float *control_inputs, float *control_cumtime,
                           float *predicted_locations);

/* Command lookup at `elapsed[b]` seconds after the publish (controller.py:110-116):
 * mode 0 = TemporalCommandSelector.get_command (commands.py:8-38, the production selector),
 * mode 1 = TemporalCommandInterpolator.get_command (commands.py:41-99).
 * cum_time[B,n], commands[B,n,2] (rows = (v, delta)), out[B,2]; indices (may be NULL) [B,2] = the command rows used.
 * _f32 works on the float32 shared-memory views in float32, _f64 on float64 arrays (the reference's test vectors). */
#define ACMPC_COMMAND_SELECT 0
#define ACMPC_COMMAND_INTERPOLATE 1
int32_t acmpc_select_commands_f32_host(acmpc_handle *h, int32_t B, int32_t n, const float *cum_time,
                                       const float *commands, const double *elapsed, int32_t mode, float *out,
                                       int32_t *indices);
int32_t acmpc_select_commands_f64_host(acmpc_handle *h, int32_t B, int32_t n, const double *cum_time,
                                       const double *commands, const double *elapsed, int32_t mode, double *out,
                                       int32_t *indices);

/* ---- Track side of the step ---------------------------------------------------
 * utils/load.py:30-35 remove_near_duplicate_points: keeps row 0 and every row whose distance to its predecessor IN THE
 * INPUT exceeds tol (the reference uses 0.0001), order preserved.  xy[M,2] -> out[*kept,2]; `out` needs room for M rows. */
int32_t acmpc_remove_near_duplicates_host(acmpc_handle *h, int32_t M, const double *xy, double tol, double *out,
                                          int32_t *kept);
/* Same for rows of `cols` >= 2 doubles whose first two columns are (x, y): whole rows are kept or dropped, like numpy's
 * track[is_not_duplicated] on a map with extra columns (z, width ...). */
int32_t acmpc_remove_near_duplicates_cols_host(acmpc_handle *h, int32_t M, int32_t cols, const double *rows, double tol,
                                               double *out, int32_t *kept);

/* perception/utils.py:107-119 smooth_track_with_polyfit(track, num_points, degree) for B tracks at once.
 * Track b = rows offsets[b] .. offsets[b+1] of points[.,2] (x, y); offsets[0] == 0; an empty track yields the
 * reference's stub line.  out[B,num_points,2].  degree 0..3 (the reference uses 2 and 3).
 * status[B] (may be NULL): 0 fitted, 1 empty track, 2 fewer than degree+1 distinct abscissae (numpy: RankWarning and the
 * minimum-norm solution; here: the highest degree the data determine).  start_index[B] (may be NULL) = the argmin of
 * perception/utils.py:116 on the 500-point scan. */
#define ACMPC_TRACK_OK 0
#define ACMPC_TRACK_EMPTY 1
#define ACMPC_TRACK_RANK_DEFICIENT 2
int32_t acmpc_smooth_tracks_polyfit_host(acmpc_handle *h, int32_t B, const int32_t *offsets, const double *points,
                                         int32_t num_points, int32_t degree, double *out, int32_t *status,
                                         int32_t *start_index);

/* perception/tracks.py:247-252 TrackLimitPerception._calculate_centre_track for B frames: left/right [B,N,2] (the
 * smoothed limits) -> centre [B,num_points,2] = degree-2 fit of (left + right) / 2 with 10 origin points in front. */
int32_t acmpc_centre_tracks_host(acmpc_handle *h, int32_t B, int32_t N, const double *left, const double *right,
                                 int32_t num_points, double *centre, int32_t *status);

/* SURVEY.md section 8d "instance -> get_control input" on the device: centre line [M,2] (closed loop, one point every
 * ds metres), per instance a map index, a lateral offset along the left normal and a heading offset ->
 * paths[B,H,3] = the next `lookahead` metres resampled to H points in the ego frame (x right, y forward,
 * spatial_mpc.py:186-187) + widths linspace(10, 6, H) (controller.py:264) -- the layout acmpc_solve_batch_device
 * reads.  d_offset_lat / d_offset_psi may be NULL (= 0).  _device: device pointers, asynchronous on `stream`
 * (the same stream argument as acmpc_solve_batch_device, so the two calls are ordered), no synchronisation. */
int32_t acmpc_extract_paths_device(acmpc_handle *h, int32_t M, const double *d_centreline, int32_t B,
                                   const int32_t *d_index, const double *d_offset_lat, const double *d_offset_psi,
                                   double lookahead, double ds, double *d_paths, void *stream);
int32_t acmpc_extract_paths_host(acmpc_handle *h, int32_t M, const double *centreline, int32_t B, const int32_t *index,
                                 const double *offset_lat, const double *offset_psi, double lookahead, double ds,
                                 double *paths);

#ifdef __cplusplus
}
#endif
#endif /* ACMPC_B200_H */
