/*
 * hmp_sweep_tpc.inl -- the FP32 sweep with one THREAD per candidate (included by hmp_kernels.cu inside namespace hmp).
 *
 * plan_kernel gives a candidate a whole warp: the 32 lanes stride over the objects and every lane repeats the per-step
 * scalar section (internal force, twist, saturation, limits, feasibility, MapGrid look-ups, smoothness sums, pose
 * integration) and the shuffle reductions. That is the right shape for few candidates (latency of one rollout) but at
 * 16k+ candidates per scene the redundant scalar work and the partly filled object iterations (50 dynamic objects in
 * 2 x 32 slots) are about half of all issued instructions. Here a warp rolls out 32 candidates at once: every lane owns
 * one candidate and walks ALL objects itself; the object records are read with warp-uniform shared-memory loads
 * (broadcast, no bank conflicts), nothing is reduced across lanes during the rollout, and the scalar section is issued
 * once per 32 candidates. The arithmetic (FP32 object loops and scalar section, FP64 pose and cell indexing) is the one of
 * plan_kernel<false, float, false>; only the summation order of the forces differs (sequential instead of lane-strided).
 *
 * The one cooperative part is the obstacle critic: a pose whose dilated-map look-up says the footprint must be rasterised
 * is broadcast to the warp and walked by all 32 lanes with footprint_pose(), pose after pose.
 *
 * Reference statements: the same as plan_kernel (file header of hmp_kernels.cu).
 */
#ifndef HMP_TPC_THREADS
#define HMP_TPC_THREADS 256
#endif
#ifndef HMP_TPC_UNROLL
#define HMP_TPC_UNROLL 2
#endif

#ifndef HMP_TPC_ESTRIN
#define HMP_TPC_ESTRIN 0   /* 1: the asin polynomial of the packed loops by Estrin's scheme (A/B) */
#endif
#ifndef HMP_TPC_PACKED
#define HMP_TPC_PACKED 1   /* static-object loop in packed FP32x2 arithmetic (FFMA2 / FADD2 / FMUL2 of sm_100), two objects per iteration */
#endif

__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
// Angle in [0, pi] between two vectors from the cosine c and the (non-negative) sine s of it, both for two vectors at once
// in packed arithmetic, without a division: asin of the smaller of (s, |c|) by a degree-7 minimax polynomial in z^2 on
// [0, sin(pi/4)] (max error 9e-8 in FP32 evaluation, the accuracy class of atan2_r), then the octant.
__device__ __forceinline__ float2 angle_from_cos_sin2(float2 c, float2 s) {
	const bool lo_x = s.x <= fabsf(c.x), lo_y = s.y <= fabsf(c.y);
	const float2 z = make_float2(lo_x ? s.x : c.x, lo_y ? s.y : c.y);
	const float2 q = __fmul2_rn(z, z);
#if HMP_TPC_ESTRIN
	// Estrin's scheme: the same degree-8 polynomial with a dependency depth of 4 instead of 9 (two more multiplications)
	const float2 q2 = __fmul2_rn(q, q);
	const float2 q4 = __fmul2_rn(q2, q2);
	const float2 e01 = __ffma2_rn(bc2(1.666642890e-01f), q, bc2(1.000000020e+00f));
	const float2 e23 = __ffma2_rn(bc2(4.341339492e-02f), q, bc2(7.508057529e-02f));
	const float2 e45 = __ffma2_rn(bc2(-2.460549290e-02f), q, bc2(4.040429929e-02f));
	const float2 e67 = __ffma2_rn(bc2(-1.801251085e-01f), q, bc2(1.457034517e-01f));
	const float2 e03 = __ffma2_rn(e23, q2, e01);
	const float2 e47 = __ffma2_rn(e67, q2, e45);
	const float2 e8 = __fmul2_rn(bc2(1.448995462e-01f), q4);
	float2 p = __ffma2_rn(__ffma2_rn(e8, q4, e47), q4, e03);   // (c8 q^4 * q^4 + e47) q^4 + e03
#else
	float2 p = bc2(1.448995462e-01f);
	p = __ffma2_rn(p, q, bc2(-1.801251085e-01f));
	p = __ffma2_rn(p, q, bc2(1.457034517e-01f));
	p = __ffma2_rn(p, q, bc2(-2.460549290e-02f));
	p = __ffma2_rn(p, q, bc2(4.040429929e-02f));
	p = __ffma2_rn(p, q, bc2(4.341339492e-02f));
	p = __ffma2_rn(p, q, bc2(7.508057529e-02f));
	p = __ffma2_rn(p, q, bc2(1.666642890e-01f));
	p = __ffma2_rn(p, q, bc2(1.000000020e+00f));
#endif
	float2 r = __fmul2_rn(z, p);   // asin(z), signed
	// s <= |c|: the angle is asin(s) in front (c >= 0) and pi - asin(s) behind; otherwise acos(c) = pi/2 - asin(c)
	r.x = lo_x ? ((c.x < 0.0f) ? PI_F - r.x : r.x) : 1.57079632679489662f - r.x;
	r.y = lo_y ? ((c.y < 0.0f) ? PI_F - r.y : r.y) : 1.57079632679489662f - r.y;
	return r;
}

struct PeopleMax {
	float hd, psi, ps;
};
// People critics of one pose with the literal per-step person pose (people with a yaw rate): heading disturbance
// (heading_disturbance_cost_function.cpp:69-86), personal space (personal_space_intrusion_cost_function.cpp:55-78), passing
// speed (passing_speed_cost_function.cpp:57-71). Same statements as the people loop of plan_kernel.
__device__ __noinline__ PeopleMax people_critics_generic(const DevPerson* people, int n_people, float tp, float rx, float ry, float rspeed,
                                                         float motion_dir, float sp_norm, bool do_psi, bool do_hd, bool do_ps,
                                                         float hd_neg_inv_2var_fov, float hd_inv_max_speed, float hd_dmin, float ps_min_dist,
                                                         PeopleMax m) {
#pragma unroll 1
	for (int p = 0; p < n_people; ++p) {
		const float4 a0 = reinterpret_cast<const float4*>(people)[4 * p];
		const float4 a1 = reinterpret_cast<const float4*>(people)[4 * p + 1];
		const float4 a2 = reinterpret_cast<const float4*>(people)[4 * p + 2];
		const float4 a3 = reinterpret_cast<const float4*>(people)[4 * p + 3];
		float pxp = fmaf(tp, a1.x, a0.x), pyp = fmaf(tp, a1.y, a0.y);
		float dx = rx - pxp, dy = ry - pyp;
		float dist = sqrt_nr(dx * dx + dy * dy);
		float yawp = a0.z, cp = a1.z, sp = a1.w;
		if (a0.w != 0.0f) {   // warp-uniform (a property of the person)
			yawp = wrapf(fmaf(tp, a0.w, a0.z));
			sincosf(yawp, &sp, &cp);
		}
		if (do_psi) {
			// personal_space_intrusion_cost_function.cpp:55-78 (asymmetric Gaussian, peak 1)
			float along = dx * cp + dy * sp;
			float vh = (along >= 0.0f) ? a3.x : a3.y;
			float vs = a3.z;
			float ga = vh * cp * cp + vs * sp * sp + a2.x;
			float gb = (vh - vs) * cp * sp;
			float gc = vh * sp * sp + vs * cp * cp + a2.w;
			float b1 = gb + a2.y, b2 = gb + a2.z;
			float det = ga * gc - b1 * b2;
			float q = (gc * dx * dx - (b1 + b2) * dx * dy + ga * dy * dy) / det;
			m.psi = fmaxf(m.psi, __expf(-0.5f * q));
		}
		if (do_hd) {
			// heading_disturbance_cost_function.cpp:69-86
			float v = 0.0f;
			if (!(rspeed < 1e-9f) && !(dist < 1e-9f)) {
				float dist_angle = atan2_r(dy, dx);
				float rel_loc = wrapf(dist_angle - yawp);
				float gamma_cc = wrapf(dist_angle + PI_F);
				float half = atan2_r(a3.w, dist);
				float dd = wrapf(motion_dir - gamma_cc);
				float g_dir = __expf(-(dd * dd) / (2.0f * half * half));
				float g_fov = __expf(rel_loc * rel_loc * hd_neg_inv_2var_fov);
				v = g_dir * g_fov * (rspeed * hd_inv_max_speed) * (hd_dmin / fmaxf(dist, hd_dmin));
			}
			m.hd = fmaxf(m.hd, v);
		}
		if (do_ps) {
			// passing_speed_cost_function.cpp:57-71
			float clearance = fmaxf(dist - ps_min_dist, 0.0f);
			m.ps = fmaxf(m.ps, sp_norm * __expf(-clearance));
		}
	}
	return m;
}

#ifndef HMP_TPC_STATIC_CALL
#define HMP_TPC_STATIC_CALL 0   /* 1: the packed static-object loop is an out-of-line function: the call boundary parks the caller's
                                   per-candidate state (pose, critic accumulators, amplified parameters) on the stack, so that the
                                   loop body has the whole register file and several pairs of objects in flight */
#endif
#ifndef HMP_TPC_CALL_UNROLL
#define HMP_TPC_CALL_UNROLL 2
#endif
struct StaticSum {
	float fx, fy, dmin, gmin;
};
// The packed static-object loop of sweep_tpc_kernel (same statements, see there), out of line.
__device__ __noinline__ StaticSum static_loop_packed(const float4* __restrict__ pairs, int npairs, float rxh, float ryh, float rxl, float ryl,
                                                     float yx, float yy, float yl2, float hx, float hy, float ih, float nbw_l2, float fovn_l2,
                                                     float nawq) {
	const float2 nrxh2 = bc2(-rxh), nryh2 = bc2(-ryh), nrxl2 = bc2(-rxl), nryl2 = bc2(-ryl);
	const float2 yx2 = bc2(yx), yy2 = bc2(yy), nyl2_2 = bc2(-yl2);
	const float2 hx2 = bc2(hx), hy2 = bc2(hy), nhx2 = bc2(-hx), ih2 = bc2(ih);
	const float2 nbw2 = bc2(nbw_l2), fovn2 = bc2(fovn_l2), nawq2 = bc2(nawq);
	const float2 NHALF2 = bc2(-0.5f), C15_2 = bc2(1.5f), HALF2 = bc2(0.5f);
	float2 fsx2 = make_float2(0.f, 0.f), fsy2 = make_float2(0.f, 0.f);
	float dminp = CUDART_INF_F, gmin = CUDART_INF_F;
	constexpr int CALL_UNROLL = HMP_TPC_CALL_UNROLL;
#pragma unroll CALL_UNROLL
	for (int p = 0; p < npairs; ++p) {
		const float4 a4 = pairs[2 * p], b4 = pairs[2 * p + 1];   // warp-uniform addresses: broadcast
		const float2 dx = __fadd2_rn(__fadd2_rn(make_float2(a4.x, a4.y), nrxh2), __fadd2_rn(make_float2(b4.x, b4.y), nrxl2));
		const float2 dy = __fadd2_rn(__fadd2_rn(make_float2(a4.z, a4.w), nryh2), __fadd2_rn(make_float2(b4.z, b4.w), nryl2));
		const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
		const float2 bx = __fadd2_rn(dx, yx2), by = __fadd2_rn(dy, yy2);
		const float2 b2 = __ffma2_rn(bx, bx, __fmul2_rn(by, by));
		float2 ia = make_float2(rsqrt_ftz(d2.x), rsqrt_ftz(d2.y));
		float2 ib = make_float2(rsqrt_ftz(b2.x), rsqrt_ftz(b2.y));
		ia = __fmul2_rn(ia, __ffma2_rn(__fmul2_rn(__fmul2_rn(d2, NHALF2), ia), ia, C15_2));
		ib = __fmul2_rn(ib, __ffma2_rn(__fmul2_rn(__fmul2_rn(b2, NHALF2), ib), ib, C15_2));
		const float2 dist = __fmul2_rn(d2, ia), bl = __fmul2_rn(b2, ib);
		const float2 sum = __fadd2_rn(dist, bl);
		const float2 w2 = __ffma2_rn(sum, sum, nyl2_2);
		float2 iw = make_float2(rsqrt_ftz(w2.x), rsqrt_ftz(w2.y));
		iw = __fmul2_rn(iw, __ffma2_rn(__fmul2_rn(__fmul2_rn(w2, NHALF2), iw), iw, C15_2));
		const float2 w = __fmul2_rn(__fmul2_rn(w2, iw), HALF2);
		dminp = fminf(dminp, fminf(dist.x, dist.y));
		gmin = fminf(gmin, fminf(fminf(fminf(d2.x, d2.y), fminf(b2.x, b2.y)), fminf(w2.x, w2.y)));
		const float2 ex = __ffma2_rn(bx, ib, __fmul2_rn(dx, ia));
		const float2 ey = __ffma2_rn(by, ib, __fmul2_rn(dy, ia));
		const float2 dot = __ffma2_rn(dx, hx2, __fmul2_rn(dy, hy2));
		const float2 crs = __ffma2_rn(dx, hy2, __fmul2_rn(dy, nhx2));
		const float2 ihd = __fmul2_rn(ia, ih2);
		const float2 cs2 = __fmul2_rn(dot, ihd), sn2 = __fmul2_rn(make_float2(fabsf(crs.x), fabsf(crs.y)), ihd);
		const float2 ar = angle_from_cos_sin2(cs2, sn2);
		const float2 expo = __ffma2_rn(w, nbw2, __fmul2_rn(__fmul2_rn(ar, ar), fovn2));
		const float2 e = make_float2(ex2_ftz(expo.x), ex2_ftz(expo.y));
		const float2 ng = __fmul2_rn(__fmul2_rn(nawq2, e), __fmul2_rn(sum, w));
		fsx2 = __ffma2_rn(ng, ex, fsx2);
		fsy2 = __ffma2_rn(ng, ey, fsy2);
	}
	return {fsx2.x + fsx2.y, fsy2.x + fsy2.y, dminp, gmin};
}

#ifndef HMP_TPC_DYN_UNROLL
#define HMP_TPC_DYN_UNROLL 1
#endif
#ifndef HMP_TPC_PPL_UNROLL
#define HMP_TPC_PPL_UNROLL 1
#endif
#ifndef HMP_TPC_LOCKSTEP
#define HMP_TPC_LOCKSTEP 0
#endif
#ifndef HMP_TPC_STRIDED
#define HMP_TPC_STRIDED 1
#endif
#ifndef HMP_TPC_PAIR_UNROLL
#define HMP_TPC_PAIR_UNROLL 1
#endif

#ifndef HMP_TPC_MIN_BLOCKS
#define HMP_TPC_MIN_BLOCKS (512 / HMP_TPC_THREADS)
#endif
// launch bounds of the FP64 instance (exact-parity mode): largest block and resident blocks per SM the registers are budgeted for
#ifndef HMP_F64_TPC_MAXTHREADS
#define HMP_F64_TPC_MAXTHREADS HMP_TPC_THREADS
#endif
// Two blocks of 256 threads per SM (128 registers, some spill) so that a 64k-candidate grid is ONE wave of 2048 warps on 2368
// slots: 36.4 ms (cfg2 seed 0) against 43.0 ms for the spill-free 225-register build, which holds 8 warps per SM and needs two
// rounds; 168 registers / 12 warps per SM: 52 ms (still two rounds). One static object in flight (two: 38.3 ms, four: 48 ms).
#ifndef HMP_F64_TPC_MINB
#define HMP_F64_TPC_MINB 2
#endif
#ifndef HMP_F64_TPC_UNROLL
#define HMP_F64_TPC_UNROLL 1   /* static objects in flight per thread of the FP64 instance */
#endif
// MINB = resident blocks per SM the registers are budgeted for: 2 (128 registers, 16 warps per SM) for launches that fill the
// GPU, 1 (up to 255 registers: the kernel takes ~200 and loses its spills) for launches that leave at most two warps per SM
// sub-partition anyway -- there a warp's speed is pure latency and the extra registers are free (cfg1: sweep 1.59 -> 1.52 ms).
// DEFER: the instance with the deferred obstacle critic (launched when KernelArgs.pose_scratch is set); the other instance is
// compiled without that code (its register pressure costs the few-object batched worlds of config 4 a quarter of their speed).
// RT = double: the exact-parity sweep (hmp_set_precision 1) in this layout -- object loops, FIS and the per-step scalar section in
// FP64 with the literal formulations of plan_kernel<false, double> (no packed loop, no float copies of the records); the critics
// stay FP32 on the FP64 poses as in every instance. A lane evaluates ONE candidate: no shuffle reductions, the scalar section
// once per candidate instead of 32 times per warp, and the 32 lanes of a warp meet the SAME object in the divergent literal
// FIS, from neighbouring candidates (similar geometry) instead of 32 different objects.
#ifdef HMP_TPC_MAXNREG
template <int MINB, bool DEFER, typename RT = float>
__global__ void __maxnreg__(HMP_TPC_MAXNREG) sweep_tpc_kernel(const KernelArgs A) {
#else
template <int MINB, bool DEFER, typename RT = float>
__global__ void __launch_bounds__(sizeof(RT) == 8 ? HMP_F64_TPC_MAXTHREADS : HMP_TPC_THREADS, MINB) sweep_tpc_kernel(const KernelArgs A) {
#endif
	using R = RT;
	using SC = RT;
	using TwistS = TwistT<RT>;
	constexpr bool F32 = (sizeof(RT) == 4);
	constexpr int NW = HMP_TPC_THREADS / 32;
	constexpr int TPC_UNROLL = (sizeof(RT) == 8) ? HMP_F64_TPC_UNROLL : HMP_TPC_UNROLL;   // static objects in flight per thread
	[[maybe_unused]] constexpr int TPC_PAIR_UNROLL = HMP_TPC_PAIR_UNROLL;
	constexpr int TPC_DYN_UNROLL = HMP_TPC_DYN_UNROLL, TPC_PPL_UNROLL = HMP_TPC_PPL_UNROLL;   // dynamic objects / people in flight per thread   // ... pairs of them in the packed loop
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t s_bar;
	__shared__ double s_wbest[NW];
	__shared__ int s_widx[NW];
	__shared__ unsigned int s_hv[HMP_NUM_MAPGRIDS];
	__shared__ unsigned int s_cnt[2];
	__shared__ bool s_last;
	__shared__ int s_base;

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int warp = tid >> 5;
	const int scene = blockIdx.y;
	const SmemLayout L = smem_layout(A.scene_stride, A.costmap_stride, A.costmap_in_smem);

	// ---- stage parameters + scene blob + costmap window with bulk TMA copies -----------------------
	if (tid == 0) {
		mbar_init(&s_bar, 1);
		for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) s_hv[g] = 0u;
		s_cnt[0] = s_cnt[1] = 0u;
	}
	__syncthreads();
	if (tid == 0) {
		uint32_t bytes = (uint32_t)sizeof(DevParams) + A.scene_stride + (A.costmap_in_smem ? A.costmap_stride : 0u);
		mbar_expect_tx(&s_bar, bytes);
		tma_bulk_g2s(smem + L.off_params, A.params, (uint32_t)sizeof(DevParams), &s_bar);
		tma_bulk_g2s(smem + L.off_scene, A.scenes + (size_t)scene * A.scene_stride, A.scene_stride, &s_bar);
		if (A.costmap_in_smem)
			tma_bulk_g2s(smem + L.off_costmap, A.costmaps + (size_t)scene * A.costmap_stride, A.costmap_stride, &s_bar);
	}
	mbar_wait(&s_bar, 0);

	const DevParams& P = *reinterpret_cast<const DevParams*>(smem + L.off_params);
	const unsigned char* blob = smem + L.off_scene;
	const DevScene& S = *reinterpret_cast<const DevScene*>(blob);
	HMP_CHECK(blockDim.x <= HMP_TPC_THREADS && (blockDim.x & 31) == 0, "block size of the thread-per-candidate sweep");
	HMP_CHECK(S.blob_bytes <= A.scene_stride && S.off_static >= sizeof(DevScene) &&
	              S.off_static + (size_t)max(S.n_static, S.n_static0) * sizeof(DevStatic) <= S.off_dynamic &&
	              S.off_dynamic + (size_t)max(S.n_dynamic, S.n_dynamic_later) * sizeof(DevDynamic) <= S.off_people &&
	              S.off_people + (size_t)S.n_people * sizeof(DevPerson) <= S.off_groups &&
	              S.off_groups + (size_t)S.n_groups * sizeof(DevGroup) <= S.blob_bytes,
	          "scene blob: record arrays overlap or leave the blob");
	HMP_CHECK(!A.costmap_in_smem || (size_t)P.size_x * P.size_y <= A.costmap_stride, "costmap window larger than its staged stride");
	// the packed copy of the static objects (16 bytes per object, pairs of 32) sits behind the staged layout
	HMP_CHECK(((L.total + 15u) & ~15u) + (HMP_TPC_PACKED ? (uint32_t)((max(S.n_static0, S.n_static) + 1) / 2) * 32u : 0u) <= dynamic_smem_size(),
	          "staging layout + packed static objects exceed the dynamic shared memory of the launch");
	const DevStatic* statics = reinterpret_cast<const DevStatic*>(blob + S.off_static);
	const DevDynamic* dynamics = reinterpret_cast<const DevDynamic*>(blob + S.off_dynamic);
	const DevPerson* people = reinterpret_cast<const DevPerson*>(blob + S.off_people);
	const DevGroup* groups = reinterpret_cast<const DevGroup*>(blob + S.off_groups);
	const uint8_t* cm = A.costmap_in_smem ? (const uint8_t*)(smem + L.off_costmap)
	                                      : (A.costmaps + (size_t)scene * A.costmap_stride);
	const uint8_t* dil = A.dilated ? (A.dilated + (size_t)scene * A.costmap_stride) : nullptr;
	const size_t grid_cells = (size_t)P.size_x * P.size_y;
	const float* mapgrids = A.mapgrids + (size_t)scene * HMP_NUM_MAPGRIDS * grid_cells;
	MapGeom G;
	G.ox = P.origin_x;
	G.oy = P.origin_y;
	G.res = P.resolution;
	G.inv_res = P.inv_resolution;
	G.sx = P.size_x;
	G.sy = P.size_y;

#if HMP_TPC_PACKED
	// Static objects once more, as pairs of hi / lo floats (o = hi + lo to 2^-48): pair p = objects 2p, 2p + 1 as
	// {xh0, xh1, yh0, yh1} {xl0, xl1, yl0, yl1}, so that one 16-byte broadcast load feeds the packed arithmetic of two objects
	// and the difference to the robot position needs no FP64 subtraction and no conversion
	[[maybe_unused]] const float4* pairs = reinterpret_cast<const float4*>(smem + ((L.total + 15u) & ~15u));
	if constexpr (F32) {
		float* pf = reinterpret_cast<float*>(smem + ((L.total + 15u) & ~15u));
		const int nsm = max(S.n_static0, S.n_static);
		for (int j = tid; j < nsm; j += blockDim.x) {
			const double2 o = reinterpret_cast<const double2*>(statics)[j];
			const float xh = (float)o.x, yh = (float)o.y;
			float* b = pf + (j >> 1) * 8 + (j & 1);
			b[0] = xh;
			b[2] = yh;
			b[4] = (float)(o.x - (double)xh);
			b[6] = (float)(o.y - (double)yh);
		}
	}
	__syncthreads();
#endif

	// Dynamic objects: dir_beta and speed are only ever read as floats, the velocity as both: their float copies
	// {dir_beta, speed, vx, vy} go over the last 16 bytes of the record (the double `speed` and the padding), saving four
	// conversions per object-step
	for (int k = tid; F32 && k < max(S.n_dynamic, S.n_dynamic_later); k += blockDim.x) {
		const DevDynamic o = dynamics[k];
		float4* w = reinterpret_cast<float4*>(const_cast<DevDynamic*>(dynamics + k)) + 3;
		*w = make_float4((float)o.dir_beta, (float)o.speed, (float)o.vx, (float)o.vy);
		reinterpret_cast<float2*>(const_cast<DevDynamic*>(dynamics + k))[4] = make_float2((float)o.psi0, 0.0f);   // over the double psi0
	}
	// People whose yaw does not change over the horizon (yaw rate 0: what the people tracker delivers) have a personal-space
	// Gaussian that depends on the candidate only through the SIDE the robot is on (front / rear variance): the quadratic
	// form q = A dx^2 + B dx dy + C dy^2 of both sides is worked out once per block and written over the person's record
	// (same 64 bytes: {x, y, vx, vy} {cos, sin, radius_eff, 0} {A, B, C front, 0} {A, B, C rear, 0}). One person with a yaw rate
	// keeps the literal per-step evaluation for everybody.
	__shared__ int s_yaw_rate;
	if (tid == 0) s_yaw_rate = 0;
	__syncthreads();
	for (int p = tid; p < S.n_people; p += blockDim.x)
		if (people[p].vth != 0.0f) atomicOr(&s_yaw_rate, 1);
	__syncthreads();
	const bool people_tab = (s_yaw_rate == 0);
	if (people_tab) {
		float4* pt = reinterpret_cast<float4*>(const_cast<unsigned char*>(blob + S.off_people));
		for (int p = tid; p < S.n_people; p += blockDim.x) {
			const DevPerson q = people[p];
			float abc[2][3];
#pragma unroll
			for (int side = 0; side < 2; ++side) {
				// personal_space_intrusion_cost_function.cpp:55-78: heading variance by side + side variance, rotated, + pose covariance
				const float vh = side ? q.var_rear : q.var_front, vs = q.var_side, cp = q.cos0, sp = q.sin0;
				const float ga = vh * cp * cp + vs * sp * sp + q.cxx;
				const float gb = (vh - vs) * cp * sp;
				const float gc = vh * sp * sp + vs * cp * cp + q.cyy;
				const float b1 = gb + q.cxy, b2 = gb + q.cyx;
				const float det = ga * gc - b1 * b2;
				abc[side][0] = gc / det;
				abc[side][1] = -(b1 + b2) / det;
				abc[side][2] = ga / det;
			}
			pt[4 * p + 0] = make_float4(q.x, q.y, q.vx, q.vy);
			// largest eigenvalue of either side's covariance is below this: q >= |d|^2 / lam, the bound the exact pruning uses
			const float lam = fmaxf(fmaxf(q.var_front, q.var_rear), q.var_side) + fabsf(q.cxx) + fabsf(q.cyy) + fabsf(q.cxy) + fabsf(q.cyx);
			pt[4 * p + 1] = make_float4(q.cos0, q.sin0, q.radius_eff, 0.5f * 1.4426950408889634f / lam);
			pt[4 * p + 2] = make_float4(abc[0][0], abc[0][1], abc[0][2], 0.0f);
			pt[4 * p + 3] = make_float4(abc[1][0], abc[1][1], abc[1][2], 0.0f);
		}
	}
	__syncthreads();

	const int T = P.T;
	const float dt = P.dt;
	const int n_vel = (T == 1) ? 1 : T - 1;  // velocities of the wrapped Trajectory, trajectory.h:43-103
	const float obstacle_costs = (float)grid_cells;            // MapGrid::obstacleCosts()
	const float unreachable_costs = (float)grid_cells + 1.0f;  // MapGrid::unreachableCellCosts()
	const bool ob_on = P.scale[HMP_COST_OBSTACLE] != 0.0;
	// Deferred obstacle critic (max aggregation with the dilated map): the rollout only records, per pose, the upper bound of
	// its footprint cost (the dilated-map value) in shared memory and the pose itself in a global scratch; the footprints are
	// walked after the horizon, poses in DESCENDING order of their bound, until the largest bound left cannot raise the maximum
	// any more. The critic is a maximum over the poses (and "any pose negative" -> -6), so the order is free; walking the most
	// expensive pose first prunes nearly all the others, where the in-loop test against the RUNNING maximum has to walk every
	// pose of an approach to an obstacle (obstacle_separation_cost_function.cpp:85-114; bit-identical results).
	const bool ob_defer = DEFER && ob_on && A.pose_scratch != nullptr && dil != nullptr && !P.occdist_sum;
	uint8_t* s_dmax = smem + ((L.total + 15u) & ~15u) + (HMP_TPC_PACKED ? ((A.scene_stride + 31u) & ~15u) : 0u);   // [T][blockDim.x]
	// scratch slot of this block (one per RESIDENT block: the grid of a batch has thousands of blocks, a few hundred at a time)
	__shared__ unsigned int s_slot;
	if (ob_defer) {
		if (tid == 0) {
			unsigned int sl = (unsigned int)(((size_t)scene * gridDim.x + blockIdx.x) % (size_t)A.pose_n_slots);
			while (atomicCAS(&A.pose_slots[sl], 0u, 1u) != 0u) sl = (sl + 1u == (unsigned int)A.pose_n_slots) ? 0u : sl + 1u;
			s_slot = sl;
		}
		__syncthreads();
	}
	double* pose_scr = ob_defer ? A.pose_scratch + (size_t)s_slot * (size_t)P.T * 3 * blockDim.x : nullptr;
	HMP_CHECK(!ob_defer || ((L.total + 15u) & ~15u) + (HMP_TPC_PACKED ? ((A.scene_stride + 31u) & ~15u) : 0u) + (uint32_t)P.T * blockDim.x <=
	                           dynamic_smem_size(),
	          "deferred obstacle critic: the per-pose bounds exceed the dynamic shared memory of the launch");

	unsigned int* counters = A.counters + (size_t)scene * 4;
	double tbest = -1.0;   // best of the candidates this thread has scored
	int tbest_idx = -1;
	unsigned int n_generated = 0, n_valid = 0;

	for (;;) {
		__syncthreads();
		if (tid == 0) s_base = (int)atomicAdd(&counters[0], blockDim.x);   // a block may be launched with fewer than HMP_TPC_THREADS threads
		__syncthreads();
		if (s_base >= A.n_work) break;
#if HMP_TPC_STRIDED
		// Ticket k gives warp w the 32 candidates of chunk w * n_tickets + k: the warps of a block (and so of an SM) take
		// chunks spread evenly over the whole sampling grid instead of consecutive ones. Neighbouring candidates cost about the
		// same (they differ in the innermost amplifiers), so consecutive chunks make whole blocks cheap or expensive and the
		// single wave of blocks ends with the most expensive SM; the lanes of a warp still hold consecutive candidates.
		const int nwb = (int)(blockDim.x >> 5);
		const int n_tickets = ((A.n_work + 31) / 32 + nwb - 1) / nwb;
		const int wk = (warp * n_tickets + s_base / (int)blockDim.x) * 32 + lane;
#else
		const int wk = s_base + tid;
#endif
		const bool active = wk < A.n_work;
		const int cand = active ? wk + A.cand_offset : 0;

		// ---- SampleAmplifierSet of this candidate (social_trajectory_generator.cpp:166-217) ----------
		float v_des, An, Bn, Cn, Ap, Bp, Cp, Aw, Bw;
		R As;   // :641: the human-action amplifier is not truncated to float; the FP32 instance rounds it
		{
			double amp[HMP_NUM_AMPLIFIERS];
			if (cand < P.n_grid) {
				int rem = cand;
#pragma unroll
				for (int a = HMP_NUM_AMPLIFIERS - 1; a >= 0; --a) {
					int n = P.amp_n[a];
					int q = rem / n;
					HMP_CHECK(n >= 1 && n <= HMP_MAX_AMP_VALUES, "amplifier axis length");
					amp[a] = __ldg(&A.amp_values[a * HMP_MAX_AMP_VALUES + (rem - q * n)]);
					rem = q;
				}
			} else {
#pragma unroll
				for (int a = 0; a < HMP_NUM_AMPLIFIERS; ++a)
					amp[a] = __ldg(&A.extra_samples[(size_t)(cand - P.n_grid) * HMP_NUM_AMPLIFIERS + a]);
			}
			// float members of SocialForceModel, social_force_model.h:422-446
			v_des = (float)((double)P.base[0] * amp[HMP_AMP_SPEED]);
			An = (float)((double)P.base[1] * amp[HMP_AMP_AN]);
			Bn = (float)((double)P.base[2] * amp[HMP_AMP_BN]);
			Cn = (float)((double)P.base[3] * amp[HMP_AMP_CN]);
			Ap = (float)((double)P.base[4] * amp[HMP_AMP_AP]);
			Bp = (float)((double)P.base[5] * amp[HMP_AMP_BP]);
			Cp = (float)((double)P.base[6] * amp[HMP_AMP_CP]);
			Aw = (float)((double)P.base[7] * amp[HMP_AMP_AW]);
			Bw = (float)((double)P.base[8] * amp[HMP_AMP_BW]);
			As = (R)amp[HMP_AMP_AS];
		}

		// ---- rollout state (per thread = per candidate) ------------------------------------------------
		double x = S.x0, y = S.y0, th = S.yaw0;   // the pose is FP64 in every instance
		SC ux = (SC)S.u0x_d, uy = (SC)S.u0y_d, uw = (SC)S.u0w_d;
		bool rejected = false;
		SC seed_x = 0, seed_w = 0;
		bool ob_neg = false;
		int ob_best = 0;
		int ob_lb = 0;   // deferred critic: largest centre-cell cost so far, a lower bound of the final maximum
		float ob_sum = 0.0f;
		float mg_last[HMP_NUM_MAPGRIDS], mg_hv[HMP_NUM_MAPGRIDS];
		int mg_codes = 0;   // 8 bits per grid: 0 ok, else -code
#pragma unroll
		for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) mg_last[g] = mg_hv[g] = 0.0f;
		float ttc_min = CUDART_INF_F;
		int ttc_first = 0x7fffffff;
		float hd_max = -CUDART_INF_F, psi_max = -CUDART_INF_F, ps_max = -CUDART_INF_F, fsi_max = -CUDART_INF_F;
		float un_x = 0.f, un_y = 0.f, un_xy = 0.f;
		int un_n = 0;
		float hcs = 0.f, vsm_x = 0.f, vsm_y = 0.f;
		TwistS prev_tw = {0, 0, 0};
		SC last_tgx = 0, last_tgy = 0;  // global velocity of the last wrapped-Trajectory velocity (TTC look-ahead)

		const bool forces_on = !P.disable_interaction;
		const R fovh = (R)P.fov_half_d, fovg = (R)P.fov_gauss_scale_d, fovn = (R)P.fov_neg_inv_2var_d;
		const R neg_inv_Bw = (R)-1 / (R)Bw;
		[[maybe_unused]] const R nbw_l2 = neg_inv_Bw * (R)1.4426950408889634, fovn_l2 = fovn * (R)1.4426950408889634;
		[[maybe_unused]] const R aw_g = (R)Aw * fovg;
		constexpr R L2E = (R)1.4426950408889634;
		[[maybe_unused]] const R nBn_l2 = -(R)Bn * L2E, Cn_l2 = (R)Cn * L2E, nBp_l2 = -(R)Bp * L2E, Cp_l2 = (R)Cp * L2E;

		for (int i = 0; i < T; ++i) {
			bool alive = active && !rejected;
#if HMP_TPC_LOCKSTEP
			__syncthreads();   // the warps of a block walk the horizon together: the step body streams through the instruction caches once per block-step
			if (!__any_sync(0xffffffffu, alive)) continue;
#else
			if (!__any_sync(0xffffffffu, alive)) break;   // warp-uniform
#endif
			double cd = 1.0, sd = 0.0;
			TwistS tw = {0, 0, 0};
			SC tgx_d = 0, tgy_d = 0;
			if (alive) {
				sincos(th, &sd, &cd);
				const double rxd = x - S.x0, ryd = y - S.y0;
				[[maybe_unused]] const float dpsi_f = (float)(th - S.yaw0);
				[[maybe_unused]] const double dpsi = th - S.yaw0;
				const double tnow = (double)i * P.dt_d;
				// -- derived robot data (world.cpp:20-33) --
				const SC speed_d = sqrt_s(ux * ux + uy * uy);
				const R heading_r = (speed_d <= (SC)0.01) ? (R)th : atan2_r((R)uy, (R)ux);
				// -- internal force (social_force_model.cpp:311-334) --
				SC fix, fiy;
				{
					SC dx = (SC)(S.glx_d - rxd), dy = (SC)(S.gly_d - ryd);
					SC dl = sqrt_s(dx * dx + dy * dy);
					SC inv = (dl <= (SC)1e-6) ? (SC)1 : (SC)1 / dl;
					fix = (SC)P.m_over_tau * ((SC)v_des * (dx * inv) - ux);
					fiy = (SC)P.m_over_tau * ((SC)v_des * (dy * inv) - uy);
				}
				const SC gdx = (SC)(S.gx_d - rxd), gdy = (SC)(S.gy_d - ryd);
				const SC goal_dist = sqrt_s(gdx * gdx + gdy * gdy);

				R fsx = 0, fsy = 0, fhx = 0, fhy = 0, fdx_r = 0, fdy_r = 0;
				float dmin = CUDART_INF_F;
				const R c_r = (R)cd, s_r = (R)sd, th_r = (R)th;
				// -- static objects (social_force_model.cpp:440-514): every object, this candidate --
				{
					const int ns = (i == 0) ? S.n_static0 : S.n_static;
					const R yx = ux * (R)P.dt_d, yy = uy * (R)P.dt_d;
					const R yl2 = yx * yx + yy * yy;
					auto static_body = [&](auto gaussian_tag, int j) {
						constexpr bool GAUSS = decltype(gaussian_tag)::value;
						const double2 o = reinterpret_cast<const double2*>(statics)[j];   // warp-uniform address: broadcast
						R dx = (R)(o.x - rxd), dy = (R)(o.y - ryd);
						R dist, ia;
						len_inv(dx * dx + dy * dy, dist, ia);
						dmin = fminf(dmin, (float)dist);
						R bx = -dx - yx, by = -dy - yy;
						R bl, ib;
						len_inv(bx * bx + by * by, bl, ib);
						R sum = dist + bl;
						R w = (R)0.5 * sqrt_nr(sum * sum - yl2);
						const bool valid = forces_on && (fabs(w) >= (R)1e-8) && !(dist < (R)1e-8);  // false for NaN too
						if (dist <= (R)1e-6) ia = (R)1;   // ignition Vector3::Normalize leaves near-zero vectors unscaled
						if (bl <= (R)1e-6) ib = (R)1;
						R ex = -dx * ia + bx * ib, ey = -dy * ia + by * ib;
						R arel = wrap_r(atan2_r(dy, dx) - heading_r);
						R gmag;
						if constexpr (F32 && GAUSS) {
							// Aw e^{-w/Bw} * g e^{-a^2 / (2 sigma^2)} with ONE exponential: both exponents pre-scaled by log2(e)
							gmag = aw_g * ex2_ftz(fmaf(w, nbw_l2, arel * arel * fovn_l2)) * (sum * w) * (R)0.25;
						} else if constexpr (!F32 && GAUSS && HMP_F64_FAST) {
							// the FP64 form of plan_kernel<false, double>: one exp instead of two
							gmag = ((R)Aw * fovg) * exp_r(fma(w, neg_inv_Bw, arel * arel * fovn)) * ((sum * (R)0.5) * w) * (R)0.5;
						} else {
							gmag = (R)Aw * exp_r(w * neg_inv_Bw) * ((sum / (R)2) * w) * (R)0.5;
							gmag *= fov_factor<R>(arel, GAUSS ? 0 : 1, fovh, fovg, fovn);
						}
						gmag = valid ? gmag : (R)0;
						fsx = fma(gmag, ex, fsx);
						fsy = fma(gmag, ey, fsy);
					};
					if (P.fov_method == 0) {
						int j0 = 0;
#if HMP_TPC_PACKED
#if HMP_TPC_STATIC_CALL
						if constexpr (F32) if (forces_on) {
							const int npairs = ns >> 1;
							const float rxh = (float)rxd, ryh = (float)ryd;
							const bool moving = !(speed_d <= (SC)0.01);
							const StaticSum ss = static_loop_packed(pairs, npairs, rxh, ryh, (float)(rxd - (double)rxh), (float)(ryd - (double)ryh), yx, yy,
							                                        yl2, moving ? ux : c_r, moving ? uy : s_r, moving ? (SC)1 / speed_d : (SC)1, nbw_l2,
							                                        fovn_l2, -0.25f * aw_g);
							if (ss.gmin > 1e-10f) {
								j0 = 2 * npairs;
								fsx = ss.fx;
								fsy = ss.fy;
								dmin = fminf(dmin, ss.dmin);
							}
						}
#else
						if constexpr (F32) if (forces_on) {
							// ---- two objects per iteration in packed FP32x2 arithmetic; same formulas as static_body ----
							const int npairs = ns >> 1;
							const float rxh = (float)rxd, ryh = (float)ryd;
							const float2 nrxh2 = bc2(-rxh), nryh2 = bc2(-ryh);
							const float2 nrxl2 = bc2(-(float)(rxd - (double)rxh)), nryl2 = bc2(-(float)(ryd - (double)ryh));
							const float2 yx2 = bc2(yx), yy2 = bc2(yy), nyl2_2 = bc2(-yl2);
							// heading as a vector: the velocity, or the yaw direction for a (nearly) standing robot (world.cpp:26-30);
							// the angle of an object relative to it is atan2(h x d, h . d), no wrap, and only its square is used
							const bool moving = !(speed_d <= (SC)0.01);
							const float2 hx2 = bc2(moving ? ux : c_r), hy2 = bc2(moving ? uy : s_r);
							const float2 nhx2 = make_float2(-hx2.x, -hx2.y);
							const float2 ih2 = bc2(moving ? (SC)1 / speed_d : (SC)1);   // 1 / |h|
							const float2 nbw2 = bc2(nbw_l2), fovn2 = bc2(fovn_l2), nawq2 = bc2(-0.25f * aw_g);
							const float2 NHALF2 = bc2(-0.5f), C15_2 = bc2(1.5f), HALF2 = bc2(0.5f);
							float2 fsx2 = make_float2(0.f, 0.f), fsy2 = make_float2(0.f, 0.f);
							float dminp = CUDART_INF_F, gmin = CUDART_INF_F;
#pragma unroll TPC_PAIR_UNROLL
							for (int p = 0; p < npairs; ++p) {
								const float4 a4 = pairs[2 * p], b4 = pairs[2 * p + 1];   // warp-uniform addresses: broadcast
								const float2 dx = __fadd2_rn(__fadd2_rn(make_float2(a4.x, a4.y), nrxh2), __fadd2_rn(make_float2(b4.x, b4.y), nrxl2));
								const float2 dy = __fadd2_rn(__fadd2_rn(make_float2(a4.z, a4.w), nryh2), __fadd2_rn(make_float2(b4.z, b4.w), nryl2));
								const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
								const float2 bx = __fadd2_rn(dx, yx2), by = __fadd2_rn(dy, yy2);   // minus the scalar body's (bx, by)
								const float2 b2 = __ffma2_rn(bx, bx, __fmul2_rn(by, by));
								float2 ia = make_float2(rsqrt_ftz(d2.x), rsqrt_ftz(d2.y));
								float2 ib = make_float2(rsqrt_ftz(b2.x), rsqrt_ftz(b2.y));
								ia = __fmul2_rn(ia, __ffma2_rn(__fmul2_rn(__fmul2_rn(d2, NHALF2), ia), ia, C15_2));   // Newton step
								ib = __fmul2_rn(ib, __ffma2_rn(__fmul2_rn(__fmul2_rn(b2, NHALF2), ib), ib, C15_2));
								const float2 dist = __fmul2_rn(d2, ia), bl = __fmul2_rn(b2, ib);
								const float2 sum = __fadd2_rn(dist, bl);
								const float2 w2 = __ffma2_rn(sum, sum, nyl2_2);
								float2 iw = make_float2(rsqrt_ftz(w2.x), rsqrt_ftz(w2.y));
								iw = __fmul2_rn(iw, __ffma2_rn(__fmul2_rn(__fmul2_rn(w2, NHALF2), iw), iw, C15_2));
								const float2 w = __fmul2_rn(__fmul2_rn(w2, iw), HALF2);
								dminp = fminf(dminp, fminf(dist.x, dist.y));
								// degenerate geometry (a zero-length vector, w ~ 0) is left to the scalar body: see below
								gmin = fminf(gmin, fminf(fminf(fminf(d2.x, d2.y), fminf(b2.x, b2.y)), fminf(w2.x, w2.y)));
								const float2 ex = __ffma2_rn(bx, ib, __fmul2_rn(dx, ia));   // minus the scalar body's (ex, ey)
								const float2 ey = __ffma2_rn(by, ib, __fmul2_rn(dy, ia));
								const float2 dot = __ffma2_rn(dx, hx2, __fmul2_rn(dy, hy2));
								const float2 crs = __ffma2_rn(dx, hy2, __fmul2_rn(dy, nhx2));
								const float2 ihd = __fmul2_rn(ia, ih2);
								const float2 cs2 = __fmul2_rn(dot, ihd), sn2 = __fmul2_rn(make_float2(fabsf(crs.x), fabsf(crs.y)), ihd);
								const float2 ar = angle_from_cos_sin2(cs2, sn2);
								const float2 expo = __ffma2_rn(w, nbw2, __fmul2_rn(__fmul2_rn(ar, ar), fovn2));
								const float2 e = make_float2(ex2_ftz(expo.x), ex2_ftz(expo.y));
								const float2 ng = __fmul2_rn(__fmul2_rn(nawq2, e), __fmul2_rn(sum, w));   // -gmag
								fsx2 = __ffma2_rn(ng, ex, fsx2);
								fsy2 = __ffma2_rn(ng, ey, fsy2);
							}
							// every guard of the scalar body passes trivially when all squared lengths are above 1e-10; otherwise (or
							// for NaN) this step's objects are done again by the scalar body
							if (gmin > 1e-10f) {
								j0 = 2 * npairs;
								fsx = fsx2.x + fsx2.y;
								fsy = fsy2.x + fsy2.y;
								dmin = fminf(dmin, dminp);
							}
						}
#endif
#endif
#pragma unroll TPC_UNROLL
						for (int j = j0; j < ns; ++j) static_body(std::true_type{}, j);
					} else {
#pragma unroll 1
						for (int j = 0; j < ns; ++j) static_body(std::false_type{}, j);
					}
				}
				// -- dynamic objects (social_force_model.cpp:338-436) + fuzzy human-action force --
				{
					const int nd = (i == 0) ? S.n_dynamic : S.n_dynamic_later;
					const R nine = (R)9 * Cst<R>::deg();
					const R speed_r = speed_d;
					if constexpr (!F32) {
						// FP64: the literal loop of plan_kernel<false, double> (double records, IEEE-class routines, literal FIS)
#pragma unroll 1
						for (int k = 0; k < nd; ++k) {
							const DevDynamic& o = dynamics[k];
							R dx = (R)(fma(tnow, o.vx, o.d0x) - rxd), dy = (R)(fma(tnow, o.vy, o.d0y) - ryd);
							R dist = sqrt_nr(dx * dx + dy * dy);
							dmin = fminf(dmin, (float)dist);
							if (!forces_on) continue;
							R angle_d = atan2_r(dy, dx);
							R rel = angle_d - (R)wrapd(o.psi0 + dpsi);
							R arel = fabs(rel);
							R side = (arel <= nine || arel >= Cst<R>::pi() - nine) ? (R)0 : ((rel <= (R)0) ? (R)-1 : (R)1);
							R rel_loc = wrap_r(rel);
							if (dist <= (R)7.5) {
								R vrx = (R)o.vx - (R)ux, vry = (R)o.vy - (R)uy;
								R vrel = sqrt_nr(vrx * vrx + vry * vry);
								if (vrel >= (R)1e-6) {
									R fov = fov_factor<R>(rel_loc, P.fov_method, fovh, fovg, fovn);
									R thab = wrap_r(th_r - angle_d);
									R en = (R)An * exp_r(div_r(-(R)Bn * thab * thab, vrel) - (R)Cn * dist) * fov;
									R ep = (R)Ap * exp_r(div_r(-(R)Bp * fabs(thab), vrel) - (R)Cp * dist) * fov * side;
									fdx_r += c_r * en + s_r * ep;
									fdy_r += s_r * en - c_r * ep;
								}
							}
							if (P.fis_on && dist <= (R)P.fis_range_d) {
								R strength = (exp_r(speed_r + (R)o.speed) - (R)1) * exp_r(-dist);
								R val, mu;
								fis_process<R>(heading_r, (R)o.dir_beta, rel_loc, angle_d, val, mu);
								if (mu > (R)0) {
									R ff = (R)1;
									if (P.fis_fov_method == 0 || P.fis_fov_method == 1)
										ff = fov_factor<R>(rel_loc, P.fis_fov_method, (R)P.fis_fov_half_d, (R)P.fis_gauss_scale_d,
										                   (R)P.fis_neg_inv_2var_d);
									R mag = As * mu * strength * ff;
									R sv, cv;
									sincos_r(val, &sv, &cv);
									fhx = fma(mag, cv, fhx);
									fhy = fma(mag, sv, fhy);
								}
							}
						}
					} else {
#pragma unroll TPC_DYN_UNROLL
					for (int k = 0; k < nd; ++k) {
						const DevDynamic& o = dynamics[k];
						R dx = (R)(fma(tnow, o.vx, o.d0x) - rxd), dy = (R)(fma(tnow, o.vy, o.d0y) - ryd);
						R dist = sqrt_nr(dx * dx + dy * dy);
						dmin = fminf(dmin, dist);
						if (!forces_on) continue;
						const float4 of = reinterpret_cast<const float4*>(&o)[3];   // {dir_beta, speed, vx, vy} as floats (prologue)
						const float opsi = reinterpret_cast<const float2*>(&o)[4].x;   // psi0 as a float (prologue)
						// World::computeObjectRelativeLocation, world.cpp:192-229 (un-normalised difference)
						R angle_d = atan2_r(dy, dx);
						R rel = angle_d - wrapf(opsi + dpsi_f);   // FP32 here (the FP64 wrap was 50 cycles of dependent latency per object)
						R arel = fabsf(rel);
						R side = (arel <= nine || arel >= Cst<R>::pi() - nine) ? (R)0 : ((rel <= (R)0) ? (R)-1 : (R)1);
						R rel_loc = wrapf(rel);

						if (dist <= (R)7.5) {
							R vrx = of.z - ux, vry = of.w - uy;
							R vrel, inv_vrel;
							len_inv(vrx * vrx + vry * vry, vrel, inv_vrel);
							if (vrel >= (R)1e-6) {
								R fov = (P.fov_method == 0) ? fovg * ex2_ftz(rel_loc * rel_loc * fovn_l2) : fov_factor<R>(rel_loc, 1, fovh, fovg, fovn);
								R thab = wrapf(th_r - angle_d);
								// exponentials as ex2 of pre-scaled arguments (as in the static loop), the two divisions by vrel as one reciprocal
								R en = An * ex2_ftz(((nBn_l2 * thab * thab) * inv_vrel) - Cn_l2 * dist) * fov;
								R ep = Ap * ex2_ftz(((nBp_l2 * fabsf(thab)) * inv_vrel) - Cp_l2 * dist) * fov * side;
								// n = (c, s); p = side * (s, -c)  (LEFT: n x z, RIGHT: n x -z)
								fdx_r += c_r * en + s_r * ep;
								fdy_r += s_r * en - c_r * ep;
							}
						}
						if (P.fis_on && dist <= (R)P.fis_range_d) {
							// social_conductor.cpp:37-105, :162-179
							R strength = (ex2_ftz((speed_r + of.y) * L2E) - (R)1) * ex2_ftz(-dist * L2E);
							R val, mu;
							fis_process<R>(heading_r, of.x, rel_loc, angle_d, val, mu);
							if (mu > (R)0) {
								R ff = (R)1;
								if (P.fis_fov_method == 0) ff = (R)P.fis_gauss_scale_d * ex2_ftz(rel_loc * rel_loc * ((R)P.fis_neg_inv_2var_d * L2E));
								else if (P.fis_fov_method == 1)
									ff = fov_factor<R>(rel_loc, 1, (R)P.fis_fov_half_d, (R)P.fis_gauss_scale_d, (R)P.fis_neg_inv_2var_d);
								R mag = As * mu * strength * ff;
								R sv, cv;
								__sincosf(val, &sv, &cv);   // |val| <= pi: 4e-7 absolute, below the error of the centroid itself
								fhx = fmaf(mag, cv, fhx);
								fhy = fmaf(mag, sv, fhy);
							}
						}
					}
					}   // F32
				}
				// TTC: first world index whose running minimum distance is within the collision distance
				// (ttc_cost_function.cpp:72-82)
				ttc_min = fminf(ttc_min, dmin);
				if (ttc_min <= P.ttc_collision_distance) ttc_first = min(ttc_first, i);

				const SC cs = (SC)cd, ss = (SC)sd;
				SC Fsx = fsx, Fsy = fsy, fdx = fdx_r, fdy = fdy_r;
				SC Fhx = 0, Fhy = 0;
				if (P.fis_on) {
					// rotate to the global frame, x force_factor (social_conductor.cpp:96-104)
					Fhx = (fhx * cs - fhy * ss) * (SC)P.fis_force_factor_d;
					Fhy = (fhx * ss + fhy * cs) * (SC)P.fis_force_factor_d;
				}
				// factorInForceCoefficients + applyNonlinearOperations (social_force_model.cpp:745-881)
				fix *= (SC)P.k_int;
				fiy *= (SC)P.k_int;
				Fsx *= (SC)P.k_stat;
				Fsy *= (SC)P.k_stat;
				fdx *= (SC)P.k_dyn;
				fdy *= (SC)P.k_dyn;
				if (P.filter_forces) {
					SC cx = fix + fdx + Fsx, cy = fiy + fdy + Fsy;
					SC mag = sqrt_s(cx * cx + cy * cy);
					if (mag >= (SC)P.max_force) {
						SC k = (SC)P.max_force / mag;
						fix *= k; fiy *= k; fdx *= k; fdy *= k; Fsx *= k; Fsy *= k;
					} else if (mag <= (SC)P.min_force) {
						SC ext = fabs(mag - (SC)P.min_force);
						SC inv = (mag <= (SC)1e-6) ? (SC)1 : (SC)1 / mag;
						fdx += ext * cx * inv;
						fdy += ext * cy * inv;
					}
				}
				// -- computeTwist (transformations.cpp:61-126) --
				const SC Fx = fix + fdx + Fsx + Fhx, Fy = fiy + fdy + Fsy + Fhy;
				bool has_force;
				if constexpr (F32) has_force = !((Fx * Fx + Fy * Fy) <= 1e-16f);   // |F| <= 1e-8 without the sqrt
				else has_force = !(sqrt(Fx * Fx + Fy * Fy) <= 1e-8);
				if (has_force && !(P.mass <= 1e-6)) {
					SC ax = Fx / (SC)P.mass, ay = Fy / (SC)P.mass;
					SC vv = cs * ax + ss * ay;
					const SC vcross = -ss * ax + cs * ay;
					// angle of the force relative to the yaw: polynomial atan2 of (F . e_yaw, F x e_yaw), no wrap needed
					// (the FP64 instance keeps the literal form Angle(atan2(Fy, Fx) - yaw), normalised)
					SC ang;
					if constexpr (F32) ang = atan2_r(vcross, vv);
					else ang = wrapd(atan2_r(Fy, Fx) - th);
					SC vw = vcross + (SC)P.rot_comp * ang;
					tw = saturate_velocity<SC>({vv, 0, vw}, (SC)P.max_vel_x, (SC)0, (SC)P.max_vel_x, (SC)P.max_vel_theta, (SC)P.back_max);
				}
				// -- adjustTwistWithAccAndGoalLimits (transformations.cpp:257-317 -> :199-255) --
				{
					TwistS vl = {ux * cs + uy * ss, 0, uw};  // computeVelocityLocal, non-holonomic
					SC smax = sqrt_s((SC)2 * (SC)P.acc_decel * goal_dist);
					SC ca = 1, sa = 0;
					if (fabs(vl.x) >= (SC)1e-4 || fabs(vl.y) >= (SC)1e-4) {
						// cos / sin of atan2(cmd.y, cmd.x) without the trigonometry (atan2(0, 0) = 0 -> (1, 0))
						SC tl = sqrt_s(tw.x * tw.x + tw.y * tw.y);
						if (tl > (SC)0) {
							ca = tw.x / tl;
							sa = tw.y / tl;
						} else if (signbit(tw.x)) {
							ca = -1;   // atan2(+-0, -0) = +-pi
						}
					}
					const SC adt_x = (SC)P.acc_x * (SC)P.dt_d, adt_y = (SC)P.acc_y * (SC)P.dt_d, adt_w = (SC)P.acc_th * (SC)P.dt_d;
					SC max_x = fmax(fmin((SC)P.max_vel_x, ca * smax), (SC)P.min_vel_x);
					SC max_y = fmax(fmin((SC)P.max_vel_y, sa * smax), (SC)P.min_vel_y);
					SC lo_x = fmax((SC)P.min_vel_x, vl.x - adt_x), hi_x = fmin(max_x, vl.x + adt_x);
					SC lo_y = fmax((SC)P.min_vel_y, vl.y - adt_y), hi_y = fmin(max_y, vl.y + adt_y);
					SC lo_w = fmax(-(SC)P.max_vel_theta, vl.w - adt_w), hi_w = fmin((SC)P.max_vel_theta, vl.w + adt_w);
					if (!P.maintain_rate) {
						tw.x = fmin(fmax(lo_x, tw.x), hi_x);
						tw.y = fmin(fmax(lo_y, tw.y), hi_y);
						tw.w = fmin(fmax(lo_w, tw.w), hi_w);
					} else {
						tw = adjust_proportional<SC>(vl, tw, lo_x, lo_y, lo_w, hi_x, hi_y, hi_w);
					}
				}
				// -- areVelocityLimitsFulfilled (social_trajectory_generator.cpp:556-582) --
				{
					SC sl = sqrt_s(tw.x * tw.x + tw.y * tw.y);
					bool trans_wrong = (P.min_vel_trans >= 0.0) && ((sl + (SC)1e-4) < (SC)P.min_vel_trans);
					bool theta_wrong = (P.min_vel_theta >= 0.0) && ((fabs(tw.w) + (SC)1e-4) < (SC)P.min_vel_theta);
					if ((trans_wrong && theta_wrong) || ((P.max_vel_trans >= 0.0) && ((sl - (SC)1e-4) > (SC)P.max_vel_trans))) {
						rejected = true;
						alive = false;
					}
				}
				if (alive) {
					if (i == 0) {
						seed_x = tw.x;
						seed_w = tw.w;
					}
					tgx_d = tw.x * cs - tw.y * ss;   // computeVelocityGlobal
					tgy_d = tw.x * ss + tw.y * cs;
				}
			}

			// =============================== critics on pose i ==========================================
			// ObstacleSeparationCostFunction (obstacle_separation_cost_function.cpp:85-114): the dilated-map test is per
			// thread, the poses that must be rasterised are walked one after the other by the whole warp
			if (ob_defer) {
				if (alive) {
					// bound of this pose's footprint cost; 255 (walk in any case) for a centre off the map. A pose whose bound
					// does not exceed the centre-cell cost of an earlier pose (a lower bound of the final maximum) and holds
					// nothing lethal / unknown can never matter: recorded as 0 = never walked.
					int dm = 255, mx, my;
					if (world_to_map(G, x, y, mx, my)) {
						HMP_CHECK((size_t)(my * G.sx + mx) < grid_cells, "dilated-map look-up outside the map");
						dm = (int)__ldg(&dil[my * G.sx + mx]);
						if (dm <= ob_lb && dm < 254) dm = 0;
						ob_lb = max(ob_lb, (int)cm[my * G.sx + mx]);
					}
					s_dmax[i * blockDim.x + tid] = (uint8_t)dm;
					if (dm != 0) {
						double* ps = pose_scr + (size_t)i * 3 * blockDim.x + tid;
						ps[0] = x;
						ps[blockDim.x] = y;
						ps[2 * blockDim.x] = th;
					}
				}
			} else {
				bool need = alive && ob_on && !ob_neg;
				if (need && dil != nullptr) {
					int mx, my;
					if (world_to_map(G, x, y, mx, my)) {
						HMP_CHECK((size_t)(my * G.sx + mx) < grid_cells, "dilated-map look-up outside the map");
						const int dmax = (int)__ldg(&dil[my * G.sx + mx]);
						// dmax < 254: no lethal / unknown cell in reach (see plan_kernel)
						const bool skip = P.occdist_sum ? (dmax == 0) : (dmax <= ob_best && dmax < 254);
						need = !skip;
					}
				}
				unsigned m = __ballot_sync(0xffffffffu, need);
				while (m) {
					const int src = __ffs(m) - 1;
					m &= m - 1;
					const double bx = __shfl_sync(0xffffffffu, x, src), by = __shfl_sync(0xffffffffu, y, src);
					const double bc = __shfl_sync(0xffffffffu, cd, src), bs = __shfl_sync(0xffffffffu, sd, src);
					bool neg = false;
					int best = 0;
					footprint_pose(P, G, cm, bx, by, bc, bs, lane, neg, best);
					const bool any_neg = __any_sync(0xffffffffu, neg);  // the first negative pose aborts the critic (-6)
					best = __reduce_max_sync(0xffffffffu, best);
					if (lane == src) {
						ob_neg = any_neg;
						if (P.occdist_sum) ob_sum += (float)best;
						else ob_best = max(ob_best, best);
					}
				}
			}
			if (alive) {
				const float rx = (float)(x - S.x0), ry = (float)(y - S.y0);
				// a negative obstacle cost aborts the scoring of this trajectory (SimpleScoredSamplingPlanner): the remaining
				// critics are never evaluated by the reference, only the rollout continues (the generator may still reject it)
				const bool dead = ob_neg && ob_on;
				// MapGridCostFunction x4 (map_grid_cost_function.cpp:142-196, :81-140)
				if (!dead) {
#pragma unroll
					for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) {
						if ((mg_codes >> (8 * g)) & 0xff) continue;
						double px = x, py = y;
						if (P.mg_xshift[g] != 0.0) {
							px += P.mg_xshift[g] * cd;
							py += P.mg_xshift[g] * sd;
						}
						if (P.mg_yshift[g] != 0.0) {
							px += P.mg_yshift[g] * (-sd);
							py += P.mg_yshift[g] * cd;
						}
						int mx, my;
						if (!world_to_map(G, px, py, mx, my)) {
							mg_codes |= 4 << (8 * g);
						} else {
							HMP_CHECK((size_t)my * P.size_x + mx < grid_cells, "MapGrid look-up outside the grid");
							float v = __ldg(&mapgrids[(size_t)g * grid_cells + (size_t)my * P.size_x + mx]);
							if (v != unreachable_costs || P.mg_kernel[g] <= 0) {
								if (v != obstacle_costs) mg_hv[g] = fmaxf(mg_hv[g], v);
							} else {
								// the neighbourhood list starts with the unreachable cell itself (see plan_kernel)
								v = (float)S.hv_prev[g];
							}
							if (P.mg_stop_on_failure[g]) {
								if (v == obstacle_costs) mg_codes |= 3 << (8 * g);
								else if (v == unreachable_costs) mg_codes |= 2 << (8 * g);
							}
							mg_last[g] = v;
						}
					}
				}
				// velocity-based critics use velocity i of the wrapped Trajectory (exists for i == 0 or i <= T - 2)
				if (i < n_vel) {
					last_tgx = tgx_d;
					last_tgy = tgy_d;
					// the critics are FP32 on float copies of the twist in every instance (as in plan_kernel)
					const float twx = (float)tw.x, twy = (float)tw.y, tww = (float)tw.w;
					const float tgx = (float)tgx_d, tgy = (float)tgy_d;
					// UnsaturatedTranslationCostFunction (:31-87)
					if (i == 0 || P.unsat_whole) {
						un_x += fabsf(twx - P.unsat_max_x);
						un_y += fabsf(twy - P.unsat_max_y);
						un_xy += fabsf(hypotf(twx, twy) - P.unsat_max_trans);
						un_n++;
					}
					// HeadingChangeSmoothness (:15-43), VelocitySmoothness (:18-50)
					if (i == 0) {
						hcs = fabsf(tww - S.vlw);
						vsm_x = fabsf(twx - S.vlx);
						vsm_y = fabsf(twy - S.vly);
					} else {
						hcs += (float)fabs(tw.w - prev_tw.w) / dt;
						vsm_x += (float)fabs(tw.x - prev_tw.x);
						vsm_y += (float)fabs(tw.y - prev_tw.y);
					}
					prev_tw = tw;
					// people critics: heading disturbance, personal space, passing speed
					const bool do_hd = (i == 0 || P.hd_whole) && P.scale[HMP_COST_HEADING_DIST] != 0.0;
					const bool do_psi = (i == 0 || P.psi_whole) && P.scale[HMP_COST_PERSONAL_SPACE] != 0.0;
					const bool do_ps = (i == 0 || P.ps_whole) && P.scale[HMP_COST_PASSING_SPEED] != 0.0;
					if (!dead && (do_hd || do_psi || do_ps)) {
						const float tp = (float)i * P.people_dt;
						const float rspeed = hypotf(tgx, tgy);
						const float motion_dir = atan2_r(tgy, tgx);
						const float sp_norm = fminf(fmaxf(rspeed * P.ps_inv_max_speed, 0.0f), 1.0f);
						if (people_tab) {
							// table path (all yaw rates zero). The heading-disturbance critic needs two angles only as squares: the
							// robot relative to the person's heading and the robot's motion direction relative to the direction
							// robot -> person; both come from their cosine and |sine| (dot / cross products over the distance) in ONE
							// packed evaluation, instead of two atan2 and three wraps
							const bool hd_ok = do_hd && !(rspeed < 1e-9f);
							const float inv_rs = hd_ok ? 1.0f / rspeed : 0.0f;
							const float mxu = tgx * inv_rs, myu = tgy * inv_rs;   // unit motion direction
							const float hd_speed = rspeed * P.hd_inv_max_speed;
							const float4* pt = reinterpret_cast<const float4*>(people);
#pragma unroll TPC_PPL_UNROLL
							for (int p = 0; p < S.n_people; ++p) {
								const float4 t0 = pt[4 * p], t1 = pt[4 * p + 1];   // warp-uniform addresses: broadcast
								const float dx = rx - fmaf(tp, t0.z, t0.x), dy = ry - fmaf(tp, t0.w, t0.y);
								float dist, inv;
								len_inv(dx * dx + dy * dy, dist, inv);
								// Exact pruning (both critics are maxima over people and poses): the personal-space value is at most
								// exp(-|d|^2 / (2 lam)) and the heading disturbance at most speed x min(1, dmin / |d|); a person whose
								// bounds (with 1e-4 of slack for the FP32 evaluation) do not exceed the running maxima cannot change
								// them. The lanes of a warp are neighbours in the sampling grid, so they mostly agree and the warp skips.
								const float d2p = dx * dx + dy * dy;
								const bool psi_live = do_psi && (ex2_ftz(-d2p * t1.w) * 1.0001f > psi_max || A.no_prune);
								const float hd_bound = hd_speed * fminf(1.0f, P.hd_dmin * inv);
								const bool hd_live = do_hd && (hd_bound * 1.0001f > hd_max || !(hd_max >= 0.0f) || A.no_prune);
								if (do_ps) {
									const float clearance = fmaxf(dist - P.ps_min_dist, 0.0f);
									ps_max = fmaxf(ps_max, sp_norm * __expf(-clearance));
								}
								if (!(psi_live || hd_live)) continue;
								const float along = dx * t1.x + dy * t1.y;
								if (psi_live) {
									const float4 tq = (along >= 0.0f) ? pt[4 * p + 2] : pt[4 * p + 3];
									const float q = tq.x * dx * dx + tq.y * dx * dy + tq.z * dy * dy;
									psi_max = fmaxf(psi_max, __expf(-0.5f * q));
								}
								if (hd_live) {
									float v = 0.0f;
									if (hd_ok && !(dist < 1e-9f)) {
										const float crs = t1.x * dy - t1.y * dx;
										const float dotm = -(mxu * dx + myu * dy), crsm = mxu * dy - myu * dx;
										const float2 ang = angle_from_cos_sin2(__fmul2_rn(make_float2(along, dotm), bc2(inv)),
										                                       __fmul2_rn(make_float2(fabsf(crs), fabsf(crsm)), bc2(inv)));
										const float half = atan2_r(t1.z, dist);
										const float g_dir = __expf(-0.5f * (ang.y * ang.y) * rcp_ftz(half * half));
										const float g_fov = __expf(ang.x * ang.x * P.hd_neg_inv_2var_fov);
										v = g_dir * g_fov * hd_speed * fminf(1.0f, P.hd_dmin * inv);
									}
									hd_max = fmaxf(hd_max, v);
								}
							}
						} else {
							// somebody turns: the literal per-step evaluation, out of line (rare path, keeps its registers out of the sweep)
							const PeopleMax pm = people_critics_generic(people, S.n_people, tp, rx, ry, rspeed, motion_dir, sp_norm, do_psi, do_hd,
							                                            do_ps, P.hd_neg_inv_2var_fov, P.hd_inv_max_speed, P.hd_dmin, P.ps_min_dist,
							                                            PeopleMax{hd_max, psi_max, ps_max});
							hd_max = pm.hd;
							psi_max = pm.psi;
							ps_max = pm.ps;
						}
					}
				}
				// FformationSpaceIntrusion (:39-78): every pose
				if (!dead && (i == 0 || P.fsi_whole) && P.scale[HMP_COST_FFORMATION] != 0.0) {
#pragma unroll 1
					for (int gidx = 0; gidx < S.n_groups; ++gidx) {
						const float4 g0 = reinterpret_cast<const float4*>(groups)[2 * gidx];
						const float ic = groups[gidx].ic;
						float dx = rx - g0.x, dy = ry - g0.y;
						float q = g0.z * dx * dx + 2.0f * g0.w * dx * dy + ic * dy * dy;
						fsi_max = fmaxf(fsi_max, __expf(-0.5f * q));
					}
				}
				// -- World::predict (world.cpp:86-114): integrate the centroid in FP64 --
				x += tgx_d * P.dt_d;
				y += tgy_d * P.dt_d;
				th = wrapd(th + tw.w * P.dt_d);
				ux = tgx_d;
				uy = tgy_d;
				uw = tw.w;
			}
		}

		// ---- deferred obstacle critic: walk the recorded poses, largest bound first ----------------------
		if (ob_defer) {
			bool more = active && !rejected;   // a rejected sample is never scored; the others hold all T poses
			for (;;) {
				int sel = -1;
				if (more) {
					int selv = 0;
					for (int i = 0; i < T; ++i) {
						const int v = (int)s_dmax[i * blockDim.x + tid];
						if (v > selv) {
							selv = v;
							sel = i;
						}
					}
					// the same test as in the loop: nothing above the maximum so far and nothing lethal / unknown in reach
					if (!(selv > ob_best || selv >= 254)) sel = -1;
					if (sel < 0) more = false;
				}
				unsigned m = __ballot_sync(0xffffffffu, sel >= 0);
				if (!m) break;
				double px = 0.0, py = 0.0, pc = 1.0, ps = 0.0;
				if (sel >= 0) {
					const double* pp = pose_scr + (size_t)sel * 3 * blockDim.x + tid;
					px = pp[0];
					py = pp[blockDim.x];
					sincos(pp[2 * blockDim.x], &ps, &pc);   // the rollout's own (cos, sin) of this yaw, bit for bit
					s_dmax[sel * blockDim.x + tid] = 0;
				}
				while (m) {
					const int src = __ffs(m) - 1;
					m &= m - 1;
					const double bx = __shfl_sync(0xffffffffu, px, src), by = __shfl_sync(0xffffffffu, py, src);
					const double bc = __shfl_sync(0xffffffffu, pc, src), bs = __shfl_sync(0xffffffffu, ps, src);
					bool neg = false;
					int best = 0;
					footprint_pose(P, G, cm, bx, by, bc, bs, lane, neg, best);
					const bool any_neg = __any_sync(0xffffffffu, neg);
					best = __reduce_max_sync(0xffffffffu, best);
					if (lane == src) {
						ob_neg = any_neg;
						ob_best = max(ob_best, best);
					}
				}
				if (ob_neg) more = false;   // -6 whatever the other poses hold
			}
		}

		// ---- TTC look-ahead (ttc_cost_function.cpp:96-134): constant-velocity continuation -------------
		if (active && !rejected && P.scale[HMP_COST_TTC] != 0.0) {
			const int n_main = 1 + n_vel;  // worlds built by World::predict(Trajectory)
			const int n_post = (n_main - T) + max(P.n_ttc_extra - 1, 0);
			// pose of world T - 1 is (x, y) minus the last integration step
			double bx = x - ux * P.dt_d, by = y - uy * P.dt_d;
			for (int j = 1; j <= n_post; ++j) {
				double rxd = bx + (SC)last_tgx * P.dt_d * j - S.x0;
				double ryd = by + (SC)last_tgy * P.dt_d * j - S.y0;
				double tnow = (double)(T - 1 + j) * P.dt_d;
				float dmin = CUDART_INF_F;
				for (int jj = 0; jj < S.n_static; ++jj) {
					const double2 o = reinterpret_cast<const double2*>(statics)[jj];
					float dx = (float)(o.x - rxd), dy = (float)(o.y - ryd);
					dmin = fminf(dmin, sqrtf(dx * dx + dy * dy));
				}
				for (int k = 0; k < S.n_dynamic_later; ++k) {
					const DevDynamic& o = dynamics[k];
					double dx = fma(tnow, o.vx, o.d0x) - rxd, dy = fma(tnow, o.vy, o.d0y) - ryd;
					dmin = fminf(dmin, (float)sqrt(dx * dx + dy * dy));
				}
				ttc_min = fminf(ttc_min, dmin);
				// world T - 1 + j is checked with timestamp (T + j) * dt when it comes from the look-ahead loop
				int stamp = (j <= n_main - T) ? (T - 1 + j) : (T + j);
				if (ttc_min <= P.ttc_collision_distance) ttc_first = min(ttc_first, stamp);
			}
		}

		// ---- the weighted total (SimpleScoredSamplingPlanner) -------------------------------------------
		double total = -1.0;
		int n_eval_grids = 0;  // MapGrid critics actually evaluated (for highest_valid_cost_)
		double hv_pre[HMP_NUM_MAPGRIDS] = {-1.0, -1.0, -1.0, -1.0};   // -1: scoring does not reach the critic
		if (active && !rejected) {
			n_generated++;
			double raw[HMP_NUM_COSTS];
			raw[HMP_COST_OBSTACLE] = (P.n_footprint == 0) ? -9.0 : (ob_neg ? -6.0 : (P.occdist_sum ? (double)ob_sum : (double)ob_best));
#pragma unroll
			for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) {
				const int code = (mg_codes >> (8 * g)) & 0xff;
				raw[HMP_COST_PATH + g] = code ? -(double)code : (double)mg_last[g];
			}
			raw[HMP_COST_UNSATURATED] = (un_n > 0) ? (double)(fmaxf(fmaxf(un_x, un_y), un_xy) / (float)un_n) : 0.0;
			// PreferForwardCostFunction
			raw[HMP_COST_BACKWARD] = (seed_x < 0.0 || (seed_x < 0.1 && fabs(seed_w) < 0.2)) ? (double)P.backward_penalty
			                                                                                : (double)(fabs(seed_w) * 10);
			{
				double c = 0.0;
				if (ttc_first != 0x7fffffff) {
					double ttc = (double)ttc_first * P.dt_d;
					if (ttc <= 0.0) ttc = 1e-4;
					c = ((double)T * P.dt_d + P.ttc_rollout_time_d) / ttc;
				}
				raw[HMP_COST_TTC] = c;
			}
			raw[HMP_COST_HEADING_CHANGE] = (double)(hcs / (float)(n_vel + 1));
			raw[HMP_COST_VEL_SMOOTHNESS] = (double)((vsm_x + vsm_y) / (float)(n_vel + 1));
			raw[HMP_COST_HEADING_DIST] = (S.n_people > 0) ? (double)hd_max : 0.0;
			raw[HMP_COST_PERSONAL_SPACE] = (S.n_people > 0) ? (double)psi_max : 0.0;
			raw[HMP_COST_FFORMATION] = (S.n_groups > 0) ? (double)fsi_max : 0.0;
			raw[HMP_COST_PASSING_SPEED] = (S.n_people > 0) ? (double)ps_max : 0.0;
			if (A.d_costs && cand == A.debug_cand) {   // parity hook (hmp_debug_sweep_candidate)
				for (int k = 0; k < HMP_NUM_COSTS; ++k) A.d_costs[k] = raw[k];
				A.d_costs[14] = seed_x; A.d_costs[15] = seed_w; A.d_costs[16] = x; A.d_costs[17] = y; A.d_costs[18] = th;
			}

			total = 0.0;
			bool aborted = false;
#pragma unroll
			for (int k = 0; k < HMP_NUM_COSTS; ++k) {
				double sc = P.scale[k];
				if (sc == 0.0 || aborted) continue;
				if (k >= HMP_COST_PATH && k <= HMP_COST_GOAL_FRONT) {
					n_eval_grids |= 1 << (k - HMP_COST_PATH);
					hv_pre[k - HMP_COST_PATH] = total;   // partial sum scoreTrajectory holds when it reaches this critic
				}
				double cst = raw[k];
				if (cst < 0.0) {
					total = cst;
					aborted = true;
					continue;
				}
				if (cst != 0.0) cst *= sc;
				total += cst;
			}
			if (total >= 0.0) {
				n_valid++;
				if (tbest < 0.0 || total < tbest || (total == tbest && cand < tbest_idx)) {
					tbest = total;
					tbest_idx = cand;
				}
			}
		}
		// highest_valid_cost_ of the four MapGrid critics: max over the warp's candidates, one atomic per warp and grid
#pragma unroll
		for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) {
			unsigned int hvb = (((n_eval_grids >> g) & 1) && mg_hv[g] > 0.0f) ? __float_as_uint(mg_hv[g]) : 0u;
			hvb = __reduce_max_sync(0xffffffffu, hvb);   // positive floats order like their bit patterns
			if (lane == 0 && hvb) atomicMax(&s_hv[g], hvb);
		}
		HMP_CHECK(!active || (cand >= 0 && cand < P.n_candidates), "explored-totals index");
		if (active && A.totals) A.totals[(size_t)scene * P.n_candidates + cand] = total;
		if (active && A.hv_pre) {
			const size_t o = ((size_t)scene * P.n_candidates + cand) * HMP_NUM_MAPGRIDS;
			reinterpret_cast<double2*>(A.hv_pre + o)[0] = make_double2(hv_pre[0], hv_pre[1]);
			reinterpret_cast<double2*>(A.hv_pre + o)[1] = make_double2(hv_pre[2], hv_pre[3]);
			*reinterpret_cast<float4*>(A.hv_val + o) = make_float4(mg_hv[0], mg_hv[1], mg_hv[2], mg_hv[3]);
		}
	}

	// ---- selection: thread -> warp -> block argmin -> last block merges (strict '<', lowest index wins ties) ----
	{
		unsigned long long bk = (tbest_idx >= 0) ? cost_key(tbest) : ~0ull;
		int bi = tbest_idx;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, o);
			int oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (oi >= 0 && (bi < 0 || ok < bk || (ok == bk && oi < bi))) {
				bk = ok;
				bi = oi;
			}
		}
		n_generated = __reduce_add_sync(0xffffffffu, n_generated);
		n_valid = __reduce_add_sync(0xffffffffu, n_valid);
		if (lane == 0) {
			s_wbest[warp] = (bi >= 0) ? __longlong_as_double((long long)bk) : -1.0;
			s_widx[warp] = bi;
			atomicAdd(&s_cnt[0], n_generated);
			atomicAdd(&s_cnt[1], n_valid);
		}
	}
	__syncthreads();
	if (ob_defer && tid == 0) atomicExch(&A.pose_slots[s_slot], 0u);   // every thread of the block is past its last scratch access
	if (tid == 0) {
		double b = -1.0;
		int bi = -1;
		for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
			double v = s_wbest[w];
			int vi = s_widx[w];
			if (vi >= 0 && v >= 0.0 && (b < 0.0 || v < b || (v == b && vi < bi))) {
				b = v;
				bi = vi;
			}
		}
		unsigned long long* bb = A.block_best + ((size_t)scene * gridDim.x + blockIdx.x) * 2;
		HMP_CHECK(bi < P.n_candidates, "block argmin index");
		bb[0] = (bi >= 0) ? cost_key(b) : ~0ull;
		bb[1] = (unsigned long long)(long long)bi;
		atomicAdd(&counters[2], s_cnt[0]);
		atomicAdd(&counters[3], s_cnt[1]);
		for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g)
			if (s_hv[g]) atomicMax(&A.hv_out[(size_t)scene * HMP_NUM_MAPGRIDS + g], s_hv[g]);
		__threadfence();
		unsigned int done = atomicAdd(&counters[1], 1u);
		s_last = (done == gridDim.x - 1);
	}
	__syncthreads();
	if (s_last && warp == 0) {
		__threadfence();
		unsigned long long bk = ~0ull;
		long long bi = -1;
		const volatile unsigned long long* bb = A.block_best + (size_t)scene * gridDim.x * 2;
		for (int b = lane; b < (int)gridDim.x; b += 32) {
			unsigned long long k = bb[2 * b];
			long long idx = (long long)bb[2 * b + 1];
			if (idx >= 0 && (k < bk || (k == bk && idx < bi))) {
				bk = k;
				bi = idx;
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, o);
			long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (oi >= 0 && (bi < 0 || ok < bk || (ok == bk && oi < bi))) {
				bk = ok;
				bi = oi;
			}
		}
		if (lane == 0) {
			if (A.best_init) {
				// best of an earlier sweep over other candidates of the same pool (the equisampled generator's)
				const double it = A.best_init[(size_t)scene * 2];
				const long long ii = (long long)A.best_init[(size_t)scene * 2 + 1];
				if (ii >= 0 && (bi < 0 || cost_key(it) < bk || (cost_key(it) == bk && ii < bi))) {
					bk = cost_key(it);
					bi = ii;
				}
			}
			A.best_out[(size_t)scene * 2 + 0] = (bi >= 0) ? __longlong_as_double((long long)bk) : -7.0;
			A.best_out[(size_t)scene * 2 + 1] = (double)bi;
		}
	}
}
