/*
 * rc_gpu.cu -- the CUDA (sm_100a) side of ref_cuda: device memory, kernels, launch order.
 *
 * Built with  nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -lineinfo
 * (float results must match the reference's C path bit for bit: no FMA contraction,
 * IEEE division, see rc_pipeline.h).  Tensor cores are not used: nothing on this path is a
 * dense contraction (BASELINE.json north_star); the kernels are integer/byte gather-scatter
 * work bounded by HBM/L2 traffic and by latency in the per-view front end.
 *
 * One process drives one GPU.  A batch of views goes through, per chunk of views:
 *   k_begin     1 thread / view           view leaf, per-view resets
 *   k_visit     1 thread / (view, node)   root->node replay: PVS, frustum cull, sort keys
 *   k_listfaces 1 thread / (view, face)   face list R_RenderFace is called on
 *   k_faces     1 thread / (view, face)   clip + project edges, surf_t setup
 *   k_fixup     1 thread / (view, face)   cross-face r_rightenter/r_rightexit reads
 *   k_scan_sort 1 warp   / (view, line)   active edges of the line, sorted -> HBM
 *   k_scan_walk 1 thread / (view, line)   R_GenerateSpans over the sorted list -> spans
 *   k_scan_blocks 1 warp / (view, line)   spans -> 8-pixel block list
 *   k_surfprep  1 thread / (view, surf)   mip, gradients, surface-cache lookup -> build jobs
 *   k_resolve   1 thread / (view, surf)   cache slot -> pool offset
 *   k_cache     1 CTA    / job (strided)  lightmap + colormap'd texel block
 *   k_draw      1 thread / 8-pixel block  perspective texture mapping + z-buffer
 * All launches are asynchronous on one stream; counts produced on the device (jobs, blocks)
 * are consumed by grid-stride loops, so the host never waits inside a batch.
 */
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "rc_host.h"
#include "rc_pipeline.h"
#include "rc_draw2d.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	snprintf (err, errlen, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString (e_)); return RC_ERR_CUDA; } } while (0)


/* Programmatic dependent launch: a kernel launched with launch_k (..., pdl = true) may be brought onto the SMs while the
 * kernel before it on the stream is still draining; it must not touch anything that kernel (or an earlier one) wrote before
 * RC_GRID_DEP_WAIT (), which returns when the preceding grid has completed and its memory is visible.  The wait is a
 * no-op in a kernel launched the ordinary way, so every kernel of the pipeline carries it. */
#ifdef RC_PDL_EARLY
/* the dependents may be brought onto the SMs as soon as every CTA of this grid has started (they still wait, in their own
 * RC_GRID_DEP_WAIT, for this grid to complete) */
#define RC_GRID_DEP_WAIT() asm volatile ("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")
#else
#define RC_GRID_DEP_WAIT() asm volatile ("griddepcontrol.wait;" ::: "memory")
#endif

#ifdef RC_TRACE
/* debug build: first start / last end of every pipeline kernel on the device's global timer (ns), printed per step by
 * rc_gpu_render_chunk's caller -- the time line inside a CUDA graph, which events cannot see */
__device__ unsigned long long g_trace[32];
struct rc_trace_scope {
	int k;
	__device__ __forceinline__ rc_trace_scope (int kk) : k (kk)
	{
		if (threadIdx.x == 0) { unsigned long long t; asm volatile ("mov.u64 %0, %globaltimer;" : "=l"(t)); atomicMin (&g_trace[2 * k], t); }
	}
	__device__ __forceinline__ ~rc_trace_scope ()
	{
		if (threadIdx.x == 0) { unsigned long long t; asm volatile ("mov.u64 %0, %globaltimer;" : "=l"(t)); atomicMax (&g_trace[2 * k + 1], t); }
	}
};
#define RC_TRACE_SCOPE(k) rc_trace_scope trace_scope_ (k)
#else
#define RC_TRACE_SCOPE(k) do { } while (0)
#endif

template <typename... KA, typename... A>
static cudaError_t launch_k (void (*kern) (KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, A &&... args)
{
	cudaLaunchConfig_t cfg;
	cudaLaunchAttribute at[1];
	memset (&cfg, 0, sizeof(cfg));
	cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
	return cudaLaunchKernelEx (&cfg, kern, static_cast<KA> (args)...);
}

#define RC_GRAPHS 4
#define RC_RF_KEEP_STATUS 1	/* rc_gpu_render_part: the status words are not cleared */
#define RC_RF_DEFER       2	/* nothing waits, no CUDA graph */
#define RC_RF_NO_MIRROR   4	/* the mirror copy is the caller's */

struct rc_gpu_s {
	int device, width, height, numsms;
	cudaStream_t stream;
	/* world */
	rc_world_t w;			/* device pointers */
	std::vector<void *> worldallocs;
	bool haveworld;
	int skybox;			/* r_drawskybox */
	int maxedges, maxsurfs;
	int *d_sintable, *d_intsintable;
	/* surface cache (persists across batches) */
	unsigned long long *d_cachekeys; int *d_cachevals; int cacheslots;
	uint8_t *d_cachepool; long long cachecap;
	/* per-chunk buffers */
	int chunkcap;			/* views the buffers below are sized for */
	int maxchunk;			/* largest chunk the memory budget allows */
	int maxviews;			/* user cap on views in flight (0: auto) */
	int facecap, edgecap, surfcap, spancap, segcap, jobcap;
	int spans_per_line;		/* sizing heuristic, doubled on RC_STAT_SPAN_POOL */
	rc_view_t *d_views;
	rc_facerec_t *d_facelist; int *d_nfaces; uint16_t *d_facepos;
	int *d_facemark; rc_nodevisit_t *d_nodevisit; int *d_viewleaf;
	rc_edge_t *d_edges; int *d_nedges;
	rc_surf_t *d_surfs;
	rc_span_t *d_spans; rc_cell_t *d_cellspan; int2 *d_slowcells;
	int *d_segspan; rc_bnd_t *d_bnd;
	int front_256, front_csize;
	int rflags;			/* RC_RF_* of the call being enqueued (rc_gpu_render_part) */
	int use_pdl;			/* programmatic dependent launch along the main stream's chain of kernels (RC_NO_PDL switches it off) */
	rc_dlight_t *d_dlights; int dlightcap;
	uint8_t *d_warpbuf; int warpcap;
	/* alias models (uploaded blobs) and per-batch entity buffers */
	rc_aliasmodel_t *d_amodels; int *d_stverts, *d_atris; uint8_t *d_poseverts, *d_skinpixels; float *d_anormdots;
	int maxverts, maxtris;
	rc_aliasinst_t *d_alias; int aliascap;
	rc_finalvert_t *d_finalverts; int finalvertcap;
	rc_particle_t *d_particles; int particlecap;
	rc_spriteinst_t *d_sprites; int spritecap; uint8_t *d_spritepixels;
	uint8_t *d_colormaps; int colormapcap;
	unsigned long long *d_zres; size_t zrescap;
	rc_aspan_t *d_aspans; long long aspancap; rc_aspantri_t *d_aspantris; int aspantricap;
	long long aspan_need; int aspantri_need;	/* what the last entity batch asked of the span queue */
	int16_t *d_zscratch; size_t zscratchcap;
	rc_bent_t *d_bents; int bentcap; int *d_leafkey; int *d_nbsurfs; int bsurfcap;
	rc_cachejob_t *d_jobs;
	int2 *d_cacheunits; int cacheunitcap;	/* (job, first row) of the row bands k_cache works on; count in alloc[8] */
	unsigned long long *d_alloc;	/* [0] spans|segments<<32, [1] cache cursor, [2] jobs, [3] cells queued for k_slowcells, [4] full cells */
	int *d_status;
	/* host-output path */
	uint8_t *d_fb; int16_t *d_zb; size_t outcap;
	uint8_t *d_d2blob; void *d_d2pics; rc_draw2d_t *d_d2cmds; int *d_d2first; int d2cmdcap, d2firstcap;	/* 2D overlay */
	int palvalid;
	uint32_t *d_pal32; int palcap; uint32_t *h_pal32[2]; cudaEvent_t evpal[2]; int palslot;	/* pinned staging ring for the palettes */
	uint8_t *d_presentin; uint32_t *d_presentout; size_t presentcap;	/* rc_gpu_present */
	/* pipelined host output (rc_gpu_render_host_async): RC_ASYNC_DEPTH calls in flight, each with its own device planes */
	uint8_t *d_afb[RC_ASYNC_DEPTH]; int16_t *d_azb[RC_ASYNC_DEPTH]; size_t aoutcap[RC_ASYNC_DEPTH];
	cudaEvent_t adone_s[RC_ASYNC_DEPTH], adone_copy[RC_ASYNC_DEPTH]; int *astatus;	/* pinned */
	int agroups;	/* groups of a pipelined call (RC_ASYNC_GROUPS) */
	int ahead, ainflight, amode;	/* amode: slot + 1 while an asynchronous call is being enqueued */
	cudaStream_t copystream, sidestream; cudaEvent_t evgroup[8], evcopy, evfork, evjoin;
	/* batches drawn in parts: the view records of a part go through one of two pinned staging buffers, so that their
	 * upload is asynchronous and the host can set up the next part while the device draws this one */
	rc_view_t *h_viewstage[2]; size_t viewstagecap[2]; cudaEvent_t ev_viewstage[2]; int viewstage_used[2], viewstage_idx;
	rc_group_fn groupfn; void *groupuser; cudaStream_t groupstream; int ngroups;	/* RC_SetGroupCallback */
	uint8_t *mirror_fb; int16_t *mirror_z;	/* RC_SetMirrorOutput */
	int mirrorgroups;
	/* deferred mirror (RC_SetMirrorDeferred): the copy of a batch to the mirror runs on the copy stream AFTER the call's
	 * kernels and the caller's stream does not wait for it -- it overlaps the next batch; a frame buffer that still has a
	 * copy in flight makes the next call that draws into it wait, RC_MirrorFence makes a stream wait for all of them */
	int mirror_deferred;
	struct { const void *fb; cudaEvent_t ev; int valid; } mpend[4];
	cudaEvent_t evrender;
	int slow_in_ends;
	int draw_first, draw_ctas, ends_ctas;	/* experiments: RC_DRAW_FIRST, RC_DRAW_CTAS, RC_ENDS_CTAS (CTAs per SM) */
	int scan_serial;
	int front_512;			/* RC_FRONT_512: keep 512-thread front-end CTAs on many-view batches (experiment knob) */
	int scan_large;			/* the scan kernel's large instance (256 active edges per line) after an overflow of the small one */
	/* the plain device path as a CUDA graph: captured once per (world, frame) parameter set, relaunched while they stay the same */
	/* a few graphs are kept: callers rotate frame buffers (double buffering, the deferred mirror) and every buffer set is
	 * its own graph */
	struct { cudaGraphExec_t exec; rc_world_t w; rc_frame_t f; unsigned long long used; } graphs[RC_GRAPHS];
	unsigned long long graph_clock; int use_graphs;
	uint8_t *host_fb; int16_t *host_z;	/* set by rc_gpu_render_host: finished view groups are copied out while the next ones are drawn */
	/* stats */
	rc_stats_t stats;
	int profiling;
	int flush_pending;		/* D_FlushCaches requested: done on the stream of the next render */
	long long launches_total;
	cudaEvent_t ev[16];
	int last_n;
	const rc_batch_t *cur_batch;
};

/* ------------------------------------------------------------------------ kernels ----- */
/* Front end, one thread-block CLUSTER per view: everything from the view leaf
 * to the view's edge and surface records in ONE launch.  The phases are the former k_begin /
 * k_visit / k_listfaces / k_faces / k_fixup bodies separated by cluster barriers.  A view's front
 * end is ~50 K warp-instructions of branchy, dependent code: one CTA per view left most of the 148
 * SMs idle at BASELINE's batch of 64, so the host picks a cluster of 1-8 CTAs per view such that the
 * grid covers the machine; the CTAs of a cluster split every phase's items and share the view's two
 * counters through distributed shared memory (rank 0's).
 *   phase 0  reset the view's facepos / facemark / nodevisit rows, background surface
 *   phase 1  rc_visit       per node-or-leaf: root->x replay (PVS, frustum, BSP sort key)
 *   phase 2  rc_list_face   per world face: is R_RenderFace called on it
 *   phase 3  rc_render_face per listed face: clip, project, edges, surf_t
 *   phase 4  fix-up of faces that read r_rightenter/r_rightexit left by an earlier face:
 *            the search for the latest earlier writer is done by the whole warp */
#ifndef FRONT_THREADS
#define FRONT_THREADS 512
#endif
#ifndef FRONT_CTAS_PER_SM
#define FRONT_CTAS_PER_SM 1
#endif
#define FRONT_WARPS (FRONT_THREADS / 32)
#ifndef FRONT_MANY
#define FRONT_MANY 256
#endif
#ifndef FRONT_MANY_MINB
#define FRONT_MANY_MINB 3	/* CTAs per SM of the 256-thread instance (many-view batches): 85 registers */
#endif
/* edge results a CTA holds per batch of faces, and its fix-up list: sized by the CTA (shared memory per CTA bounds how many
 * views an SM has in flight in the many-view instance) */
#define FRONT_EDGE_SLOTS_OF(ft) ((ft) >= 256 ? 1024 : 384)
#define FRONT_FIXLIST_OF(ft) ((ft) >= 256 ? 1024 : 256)
#define FRONT_SMEM_OF(ft) ((sizeof(rc_edgeres_t) + sizeof(uint32_t)) * FRONT_EDGE_SLOTS_OF (ft))

/* L2 prefetch of a slice of a read-only array: the front end chases pointers through the BSP
 * (node -> plane -> child ...), and after a cold start (the bench flushes L2 between steps) every
 * level would otherwise be a dependent HBM miss */
__device__ __forceinline__ void front_prefetch (const void *p, size_t bytes)
{
	const unsigned lines = (unsigned)((bytes + 127) >> 7);
	const unsigned per = (lines + gridDim.x - 1) / gridDim.x;
	const unsigned lo = per * blockIdx.x;
	unsigned hi = lo + per;
	if (hi > lines) hi = lines;
	const char *base = reinterpret_cast<const char *>(p);
	for (unsigned i = lo + threadIdx.x; i < hi; i += blockDim.x)
		asm volatile ("prefetch.global.L2 [%0];" :: "l"(base + ((size_t)i << 7)));
}

/* FT threads per CTA: 512 when a batch's views fit the machine in one wave (all the threads a view can get), 256 --
 * two CTAs per SM -- when there are more views than SMs: the phases are latency bound, so two views per SM nearly
 * double the front end's throughput on many-view batches (160x120 x 4096: 1.03 -> 0.6 ms) */
template <int FT>
__global__ void __launch_bounds__(FT, FT == FRONT_MANY ? FRONT_MANY_MINB : 1) k_front (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	RC_GRID_DEP_WAIT ();
	RC_TRACE_SCOPE (1);
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster ();
	extern __shared__ __align__(16) unsigned char front_smem[];
	__shared__ int s_counters[3];			/* listed faces, emitted edges, BSP key of the first sky face (rank 0's copy is the live one) */
	const int csize = (int)cluster.num_blocks (), crank = (int)cluster.block_rank ();
	const int vi = blockIdx.x / csize, tid = threadIdx.x;
	const int gtid = crank * FT + tid, gthreads = csize * FT;
	int *ctr = cluster.map_shared_rank (s_counters, 0);
	if (tid < 3) s_counters[tid] = tid == 2 ? 0x7FFFFFFF : 0;
#ifdef RC_FRONT_TIMING
	long long tq[16];
	int tqn = 0;
	#define TQ() do { if (tid == 0) tq[tqn++] = clock64 (); } while (0)
#else
	#define TQ() do { } while (0)
#endif
	TQ ();
	front_prefetch (w.nodes, sizeof(rc_node_t) * (size_t)w.numnodes);
	front_prefetch (w.planes, sizeof(rc_plane_t) * (size_t)w.numplanes);
	front_prefetch (w.nodeaux, sizeof(rc_nodeaux_t) * (size_t)w.numnodes);
	front_prefetch (w.leafs, sizeof(rc_leaf_t) * (size_t)w.numleafs);
	front_prefetch (w.leafdfs, sizeof(int) * (size_t)w.numleafs);
	front_prefetch (w.marksurfaces, sizeof(int) * (size_t)w.nummarksurfaces);
	front_prefetch (w.markleaf, sizeof(int) * (size_t)w.nummarksurfaces);
	front_prefetch (w.faces, sizeof(rc_face_t) * (size_t)w.numfaces);
	front_prefetch (w.fedges, sizeof(rc_fedge_t) * (size_t)w.numsurfedges);
	{
		uint16_t *facepos = f.facepos + (size_t)vi * w.numfaces;
		int *facemark = f.facemark + (size_t)vi * w.numfaces;
		unsigned long long *nv = reinterpret_cast<unsigned long long *>(f.nodevisit + (size_t)vi * w.numnodes);
		for (int i = gtid; i < w.numfaces; i += gthreads) { facepos[i] = RC_FACEPOS_NONE; facemark[i] = 0x7F7F7F7F; }
		for (int i = gtid; i < w.numnodes; i += gthreads) nv[i] = 0ull;
		int *leafkey = f.leafkey + (size_t)vi * w.numleafs;
		for (int i = gtid; i < w.numleafs; i += gthreads) leafkey[i] = 0x7FFFFFFF;
		if (gtid == 0)
			rc_view_begin (&w, &f, vi);
	}
	cluster.sync ();
	TQ ();
	/* visit and face items are long, branchy and all different: a warp pays for the union of its
	 * lanes' paths, so items are dealt round-robin to the WARPS of the cluster (item i -> warp
	 * i % nwarps, lane i / nwarps) instead of filling one warp after the other */
	const int nwarps = csize * (FT / 32);
	const int spread = (tid & 31) * nwarps + crank * (FT / 32) + (tid >> 5);
	{
		/* phase 1a: per-node frustum / plane-side tests into shared memory (every CTA of the cluster
		 * keeps its own copy: the tests are cheap, remote shared memory is not) */
		uint16_t *s_mask = reinterpret_cast<uint16_t *>(front_smem);
		for (int i = tid; i < w.numnodes; i += FT)
		{
			s_mask[i] = rc_node_test (&w, &f, vi, i);
			asm volatile ("prefetch.global.L1 [%0];" :: "l"(&w.nodeaux[i]));	/* the walk reads these */
		}
		__syncthreads ();
		TQ ();
		/* phase 1b: root->x replay over the masks */
		for (int x = spread; x < w.numnodes + w.numleafs; x += gthreads)
			rc_visit_walk (&w, &f, vi, x, s_mask);
	}
	cluster.sync ();
	TQ ();
	/* phase 1c: visited leaves stamp their faces */
	if (w.marksplit)
	{
		for (int m = gtid; m < w.nummarksurfaces; m += gthreads)
			rc_mark_face (&w, &f, vi, m);
		cluster.sync ();
	}
	TQ ();
	for (int fi = gtid; fi < w.numworldfaces; fi += gthreads)
		rc_list_face (&w, &f, vi, fi, &ctr[0], &ctr[2]);
	cluster.sync ();
	TQ ();
	if (w.drawskybox && ctr[2] != 0x7FFFFFFF)
	{
		/* a sky face was reached: the six faces of the sky box join the list (R_EmitSkyBox) */
		const int listed = ctr[0];
		cluster.sync ();
		if (gtid < 6)
			rc_list_skybox (&w, &f, vi, gtid, listed < f.facecap ? listed : f.facecap, ctr[2]);
		if (gtid == 0)
			ctr[0] = listed + 6;
		cluster.sync ();
	}
	const int nfraw = ctr[0];
	const int nf = nfraw < f.facecap ? nfraw : f.facecap;
	{
		/* phase 3, two steps per batch of faces (this CTA's contiguous share of the list):
		 *   a  one thread per polygon EDGE: edge-cache decision, clip, project   (rc_face_edge)
		 *   b  one thread per face: fold the edge results, surf_t, edge records  (rc_face_finish)
		 * A face is a serial chain of 4-8 edges in the reference; giving every edge its own thread cuts the
		 * phase's critical path by that factor.  Edge results live in shared memory. */
		rc_edgeres_t *s_res = reinterpret_cast<rc_edgeres_t *>(front_smem);
		constexpr int FRONT_EDGE_SLOTS = FRONT_EDGE_SLOTS_OF (FT);
		uint32_t *s_owner = reinterpret_cast<uint32_t *>(front_smem + sizeof(rc_edgeres_t) * FRONT_EDGE_SLOTS);
		__shared__ int s_wsum[(FT / 32)], s_included, s_total;
		const rc_facerec_t *list = f.facelist + (size_t)vi * f.facecap;
		const int plo = (int)((long long)nf * crank / csize), phi = (int)((long long)nf * (crank + 1) / csize);
		const int lane = tid & 31, wid = tid >> 5;
		for (int pb = plo; pb < phi; )
		{
			const int p = pb + tid;
			int ne = 0;
			bool toobig = false;
			if (p < phi)
			{
				ne = w.faces[list[p].face].numedges;
				if (ne > RC_MAX_FACE_EDGES) { toobig = true; ne = 0; }
			}
			/* block-wide exclusive scan of the edge counts */
			int incl = ne;
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const int t = __shfl_up_sync (0xFFFFFFFFu, incl, o);
				if (lane >= o) incl += t;
			}
			if (lane == 31) s_wsum[wid] = incl;
			if (tid == 0) { s_included = 0; s_total = 0; }
			__syncthreads ();
			int wbase = 0;
			for (int q = 0; q < wid; q++) wbase += s_wsum[q];
			const int off = wbase + incl - ne;
			const bool inc = p < phi && off + ne <= FRONT_EDGE_SLOTS;
			if (inc)
			{
				atomicAdd (&s_included, 1);
				atomicMax (&s_total, off + ne);
				for (int i = 0; i < ne; i++)
					s_owner[off + i] = (uint32_t)tid | ((uint32_t)i << 16);
			}
			__syncthreads ();
			TQ ();
			const int total = s_total, included = s_included;
			for (int e = tid; e < total; e += FT)
			{
				const uint32_t own = s_owner[e];
				rc_face_edge (&w, &f, vi, pb + (int)(own & 0xFFFFu), (int)(own >> 16), &s_res[e]);
			}
			__syncthreads ();
			TQ ();
			/* b: fold, then ONE reservation of edge records per batch (a block scan of the faces'
			 * counts; same-address atomics from every face thread serialise), then emit */
			rc_facefold_t st;
			st.neb = 0;
			if (inc && !toobig)
				rc_face_fold (&w, &f, vi, p, &s_res[off], &st);
			int einc = st.neb;
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const int t = __shfl_up_sync (0xFFFFFFFFu, einc, o);
				if (lane >= o) einc += t;
			}
			if (lane == 31) s_wsum[wid] = einc;
			__syncthreads ();
			if (tid == 0)
			{
				int tot = 0;
				for (int q = 0; q < (FT / 32); q++) tot += s_wsum[q];
				s_total = tot ? atomicAdd (&ctr[1], tot) : 0;
			}
			int ebase = 0;
			for (int q = 0; q < wid; q++) ebase += s_wsum[q];
			__syncthreads ();
			ebase += s_total + einc - st.neb;
			if (inc)
			{
				if (toobig)
				{
					rc_surf_t *surf = f.surfs + (size_t)vi * f.surfcap + (p + 2);
					surf->flags = -1; surf->hasspans = 0;
					atomicOr (f.status, RC_STAT_EDGE_OVERFLOW);
				}
				else
					rc_face_emit (&w, &f, vi, p, &s_res[off], &st, ebase);
			}
			__syncthreads ();
			pb += included;		/* included faces are a prefix of the batch (offsets grow) */
		}
	}
	cluster.sync ();
	TQ ();
	if (gtid == 0) { f.nfaces[vi] = nfraw; f.nedges[vi] = ctr[1]; }
	{
		/* phase 4: the faces that need the fix-up are few and clustered in list order: every CTA compacts them into a
		 * list and the warps of the whole cluster take them round robin (one warp-wide search per face) */
		const int lane = tid & 31, wid = tid >> 5;
		const rc_surf_t *surfs = f.surfs + (size_t)vi * f.surfcap;
		const rc_facerec_t *list = f.facelist + (size_t)vi * f.facecap;
		constexpr int FRONT_FIXLIST = FRONT_FIXLIST_OF (FT);
		__shared__ int s_nfix, s_fix[FRONT_FIXLIST];
		if (tid == 0) s_nfix = 0;
		__syncthreads ();
		for (int p = tid; p < nf; p += FT)
			if (surfs[p + 2].miplevel & 4)
			{
				const int k = atomicAdd (&s_nfix, 1);
				if (k < FRONT_FIXLIST) s_fix[k] = p;
			}
		__syncthreads ();
		const int nfix = s_nfix;
		const bool listed = nfix <= FRONT_FIXLIST;		/* else: every warp walks its own 32 faces, as a list would */
		const int nitems = listed ? nfix : ((nf + 31) & ~31);
		for (int it = crank * (FT / 32) + wid; it < nitems; it += csize * (FT / 32))
		{
			int pp;
			if (listed)
				pp = s_fix[it];
			else
			{
				pp = it;
				if (pp >= nf || !(surfs[pp + 2].miplevel & 4)) continue;
			}
			const int mykey = list[pp].key;
			long long bre = -1, brx = -1;		/* key << 32 | slot: keys are unique */
			for (int q = lane; q < nf; q += 32)
			{
				const int k = list[q].key;
				const int m = surfs[q + 2].miplevel;
				if (k < mykey)
				{
					const long long kv = ((long long)k << 32) | (unsigned)q;
					if ((m & 1) && kv > bre) bre = kv;
					if ((m & 2) && kv > brx) brx = kv;
				}
			}
			for (int o = 16; o; o >>= 1)
			{
				const long long a = __shfl_xor_sync (0xFFFFFFFFu, bre, o), b = __shfl_xor_sync (0xFFFFFFFFu, brx, o);
				if (a > bre) bre = a;
				if (b > brx) brx = b;
			}
			if (lane == 0)
				rc_face_fixup_finish (&w, &f, vi, pp, bre < 0 ? -1 : (int)(bre & 0xFFFFFFFF), brx < 0 ? -1 : (int)(brx & 0xFFFFFFFF));
		}
	}
	cluster.sync ();		/* rank 0's shared memory must outlive the other ranks' reads */
#ifdef RC_FRONT_TIMING
	TQ ();
	if (tid == 0)
		for (int i = 0; i + 1 < tqn; i++)
			atomicMax (&f.status[4 + i], (int)(tq[i + 1] - tq[i]));
#endif
}

/* Scan stage (R_ScanEdges, R_InsertNewEdges, R_StepActiveU, R_GenerateSpans), one
 * kernel, one WARP per (view, scanline), 8 lines per CTA.  Integer work throughout.
 *   A  the CTA sweeps the view's edge records once (32 B each, coalesced) and keeps in shared
 *      memory those that touch any of its 8 lines
 *   B1 each warp filters the candidates active on its line (ballot + popc compaction), u on the line
 *      is u0 + (line - v) * u_step in wrapping int32
 *   B2 rank sort on u; equal u (shared vertices, re-emitted clipped edges) are ordered by the full
 *      128-bit key that reproduces the reference's linked-list order (rc_aet_make)
 *   B3 the reference walks the sorted edges keeping a key-ordered surface stack: inherently serial.
 *      World surfaces have UNIQUE keys, so the top of the stack is simply the active surface with
 *      the smallest key.  Surfaces with a leading edge on the line are ranked by key (<= 64 of them:
 *      one bit each); every edge toggles the bits of its trailing / leading surface; an inclusive
 *      XOR scan over the sorted edges gives the active set after every edge; find-first-set gives the
 *      top.  Spans are the runs of constant top between change points, empty runs dropped.
 *      Lines where this model does not hold -- brush-entity surfaces (equal keys, 1/z tie break,
 *      r_edge.c:435-458), a surface with two leading edges, a trailing edge before / without its
 *      leading edge (spanstate going negative), more than 64 surfaces -- are walked by lane 0 with the
 *      reference's own algorithm (rc_span_walk).
 *   C  spans are published: one atomicAdd per line reserves span + checkpoint records, then the cell
 *      map the draw kernel indexes (binary search over the line's span starts). */
#ifndef SCAN_WARPS
#define SCAN_WARPS 8
#endif
#ifndef SCAN_CAND
#define SCAN_CAND  768
#endif
#define SCAN_LEADS 64
#ifndef SCAN_SMALL_AET
#define SCAN_SMALL_AET 128
#define SCAN_SMALL_CAND 384
#endif

struct rc_aet_sorted {
	const int *su; const uint32_t *ss;
	__device__ __forceinline__ int iu (int k) const { return su[k] >> 20; }
	__device__ __forceinline__ uint32_t surfs (int k) const { return ss[k]; }
	__device__ __forceinline__ int u (int k) const { return su[k]; }
};

/* AET: active edges a scanline may hold, CAND: edges that may touch the CTA's lines.  Shared memory per CTA bounds the
 * occupancy of this kernel, so the sizes are template parameters: the host starts with the small instance and moves to
 * the large one for good when a scene overflows it (RC_STAT_AET_OVERFLOW, rc_gpu_render). */
template <int AET>
struct scan_warp_t {			/* per-warp shared memory */
	static constexpr int SPANS = AET + 2;
	union {
		struct { unsigned long long ku[AET]; uint16_t kidx[AET]; } k;	/* B1/B2: unsorted keys: u (biased) << 32 | step code */
		struct { int16_t x0[SPANS], x1[SPANS]; uint16_t sf[SPANS]; } sp;	/* B3/C: change points, spans */
	} a;
	unsigned long long klo[32];	/* B1/B2, lines with <= 32 active edges: the low half of the 128-bit order key (rc_aet_make) */
	int      su[SPANS];		/* sorted: u (12.20); at publish time: first segment of every span */
	uint32_t ss[AET];		/* sorted: surfs[0] | surfs[1] << 16 */
	int      lkey[SCAN_LEADS], tcount[SCAN_LEADS];
	uint16_t lsurf[SCAN_LEADS], rsurf[SCAN_LEADS];
	uint8_t  lk[SCAN_LEADS], lrank[SCAN_LEADS];
	uint8_t  lidx[AET];
};

template <int AET, int CAND>
struct scan_cta_t {
	int      c_u[CAND], c_step[CAND];
	uint32_t c_vv[CAND], c_sf[CAND], c_rank[CAND];
	int      ncand;
	scan_warp_t<AET> warp[SCAN_WARPS];
};

template <int AET, int CAND>
__device__ __forceinline__ void scan_full_key (const scan_cta_t<AET, CAND> *sh, int c, int line, unsigned long long *hi, unsigned long long *lo)
{
	rc_edge_t ed;
	rc_aet_t a;
	ed.u = sh->c_u[c]; ed.u_step = sh->c_step[c];
	ed.v = (int16_t)(sh->c_vv[c] & 0xFFFF); ed.v2 = (int16_t)(sh->c_vv[c] >> 16);
	ed.surfs[0] = (uint16_t)(sh->c_sf[c] & 0xFFFF); ed.surfs[1] = (uint16_t)(sh->c_sf[c] >> 16);
	ed.rank = sh->c_rank[c];
	rc_aet_make (&ed, line, &a);
	*hi = a.hi; *lo = a.lo;
}

template <int AET, int CAND>
__global__ void __launch_bounds__(SCAN_WARPS * 32) k_scan (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	RC_GRID_DEP_WAIT ();
	RC_TRACE_SCOPE (2);
	extern __shared__ __align__(16) unsigned char scan_smem[];
	constexpr int SPANS = AET + 2;
	scan_cta_t<AET, CAND> *sh = reinterpret_cast<scan_cta_t<AET, CAND> *>(scan_smem);
	const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	const int vi = blockIdx.x;
	const int line0 = blockIdx.y * SCAN_WARPS, line = line0 + wid;
	const rc_view_t *vw = &f.views[vi];
	const int vy0 = vw->vrect_y, vy1 = vw->vrectbottom;
	const int cw = (f.width + 7) >> 3;

	/* ---- A: candidates of the CTA's 8 lines ---- */
	if (threadIdx.x == 0)
		sh->ncand = 0;
	__syncthreads ();
	if (line0 < vy1 && line0 + SCAN_WARPS > vy0)
	{
		const rc_edge_t *edges = f.edges + (size_t)vi * f.edgecap;
		int E = f.nedges[vi];
		if (E > f.edgecap) E = f.edgecap;
		const int lhi = line0 + SCAN_WARPS - 1;
		for (int e = threadIdx.x; e < E; e += SCAN_WARPS * 32)
		{
			const int4 q = __ldg (reinterpret_cast<const int4 *>(&edges[e]));	/* u, u_step, v|v2, surfs */
			const int v = (int16_t)(q.z & 0xFFFF), v2 = (int16_t)((unsigned)q.z >> 16);
			if (v <= lhi && v2 >= line0)
			{
				const int slot = atomicAdd (&sh->ncand, 1);
				if (slot < CAND)
				{
					sh->c_u[slot] = q.x; sh->c_step[slot] = q.y; sh->c_vv[slot] = (uint32_t)q.z; sh->c_sf[slot] = (uint32_t)q.w;
					sh->c_rank[slot] = __ldg (&edges[e].rank);
				}
			}
		}
	}
	__syncthreads ();
	int ncand = sh->ncand;
	if (ncand > CAND)
	{
		if (threadIdx.x == 0) atomicOr (f.status, RC_STAT_AET_OVERFLOW);
		ncand = CAND;
	}
	if (line >= f.height)
		return;
	int2 *cells = reinterpret_cast<int2 *>(f.cellspan + ((size_t)vi * f.height + line) * cw);
	if (line < vy0 || line >= vy1 || (vw->rdflags & 1))
	{
		for (int c = lane; c < cw; c += 32)
			cells[c] = make_int2 (-1, 0);
		if ((vw->rdflags & 1 /* RDF_NOWORLDMODEL */) && f.zb)
		{
			/* no world, no spans (r_main.c:889); the whole z-buffer is 0xFFFF for the entities (d_modech.c:103-106) */
			int16_t *zrow = f.zb + ((size_t)vi * f.height + line) * f.width;
			for (int x = lane; x < f.width; x += 32)
				zrow[x] = (int16_t)-1;
		}
		return;
	}
	scan_warp_t<AET> *sw = &sh->warp[wid];

	/* ---- B1: active edges of this line ---- */
	int n = 0;
	for (int base = 0; base < ncand; base += 32)
	{
		const int c = base + lane;
		bool act = false;
		unsigned long long key = 0ull, klo = 0ull;
		if (c < ncand)
		{
			const uint32_t vv = sh->c_vv[c];
			const int v = (int16_t)(vv & 0xFFFF), v2 = (int16_t)(vv >> 16);
			if (v <= line && v2 >= line)
			{
				/* the high word of rc_aet_make's key: u on this line, then (edges that were already
				 * active on the previous line) the larger step first; new edges before old ones */
				const uint32_t step = (uint32_t)sh->c_step[c];
				uint32_t code = 0;
				if (line > v)
				{
					code = ~(step ^ 0x80000000u);
					if (code == 0) code = 1;
				}
				act = true;
				key = ((unsigned long long)(((uint32_t)sh->c_u[c] + (uint32_t)(line - v) * step) ^ 0x80000000u) << 32) | code;
				/* the low half (rc_aet_make): later start line first, then the newedges[] order */
				const uint32_t rk = sh->c_rank[c];
				klo = ((unsigned long long)((uint32_t)(~v) & 0xFFFFu) << 32) | ((rk >> 31) ? rk : (~rk & 0x7FFFFFFFu));
			}
		}
		const unsigned m = __ballot_sync (0xFFFFFFFFu, act);
		const int pos = n + __popc (m & lt);
		if (act && pos < AET)
		{
			sw->a.k.ku[pos] = key; sw->a.k.kidx[pos] = (uint16_t)c;
			if (pos < 32) sw->klo[pos] = klo;
		}
		n += __popc (m);
	}
	if (n > AET)
	{
		if (lane == 0) atomicOr (f.status, RC_STAT_AET_OVERFLOW);
		n = AET;
	}
	__syncwarp ();

	rc_surf_t *surfs = f.surfs + (size_t)vi * f.surfcap;
	const int head_u = vw->vrect_x, tail_u = vw->vrectright;
	bool fallback = false;
	int nsp = 0;
	if (n <= 32)
	{
		/* ---- the common case, one edge per lane, everything after the sort in registers ----
		 * B2  rank sort on the 64-bit key (uniform shared-memory reads of the other keys); exact ties take the full key
		 * B3  pairing: every trailing edge finds the lane of its surface's leading edge by a bit-sliced comparison of
		 *     the 16-bit surface numbers (one ballot per bit); duplicates by match.any.  Lead k toggles bit k, its
		 *     trailing edge toggles the same bit; the inclusive XOR scan is the set of active leads after every edge,
		 *     the top of the reference's stack is the member with the smallest BSP key (unique for world surfaces).
		 *     Change points -> runs -> spans as below.  Anything the model does not cover (brush-entity surfaces, a
		 *     surface with two leading or two trailing edges, a trailing edge before / without its leading edge) is
		 *     walked by lane 0 with the reference's own algorithm. */
		const unsigned long long ui = lane < n ? sw->a.k.ku[lane] : ~0ull;
		const int myc = lane < n ? (int)sw->a.k.kidx[lane] : 0;
		/* rank by u alone (32-bit compares), then the edges that share their u -- twin edges are common: a clipped
		 * shared edge is emitted once per face -- order themselves inside their group by the rest of the key
		 * (step code, start line, emission order: unique per edge) */
		const uint32_t u32 = (uint32_t)(ui >> 32);
		const uint32_t *ku32 = reinterpret_cast<const uint32_t *>(sw->a.k.ku);
		int rank = 0;
		for (int j = 0; j < n; j++)
			rank += (ku32[2 * j + 1] < u32) ? 1 : 0;
		{
			unsigned grp = 0u;
			if (lane < n)
				grp = __match_any_sync (n == 32 ? 0xFFFFFFFFu : (1u << n) - 1u, u32) & ~(1u << lane);
			if (grp)
			{
				const uint32_t ci = (uint32_t)ui;
				const unsigned long long li = sw->klo[lane];
				for (; grp; grp &= grp - 1u)
				{
					const int j = __ffs (grp) - 1;
					const uint32_t cj = ku32[2 * j];
					rank += (cj < ci || (cj == ci && sw->klo[j] < li)) ? 1 : 0;
				}
			}
		}
		if (lane < n)
		{
			sw->su[rank] = (int)((uint32_t)(ui >> 32) ^ 0x80000000u);
			sw->ss[rank] = sh->c_sf[myc];
		}
		__syncwarp ();
		const int uk = lane < n ? sw->su[lane] : 0;
		const uint32_t sfk = lane < n ? sw->ss[lane] : 0u;
		const uint32_t s0 = sfk & 0xFFFFu, s1 = sfk >> 16;
		int key = 0x7FFFFFFF;
		if (s1)
		{
			const int4 g = *reinterpret_cast<const int4 *>(&surfs[s1]);	/* key, flags, face, insubmodel */
			key = g.x;
			fallback = g.w != 0;
		}
		sw->lkey[lane] = key;
		/* duplicates: two leading (trailing) edges of one surface */
		const unsigned dl = __match_any_sync (0xFFFFFFFFu, s1 ? s1 : 0x10000u + lane);
		const unsigned dt = __match_any_sync (0xFFFFFFFFu, s0 ? s0 : 0x10000u + lane);
		if ((dl & (dl - 1)) | (dt & (dt - 1)))
			fallback = true;
		/* lanes whose leading surface is my trailing surface */
		unsigned eq = __ballot_sync (0xFFFFFFFFu, s1 != 0);
		const unsigned anyid = __reduce_or_sync (0xFFFFFFFFu, s0 | s1);
		for (unsigned b = 0; (anyid >> b) != 0; b++)
		{
			const unsigned lb = __ballot_sync (0xFFFFFFFFu, (s1 >> b) & 1u);
			eq &= lb ^ (((s0 >> b) & 1u) - 1u);
		}
		int lpos = 0;
		if (s0)
		{
			if (eq == 0u || (int)(__ffs (eq) - 1) >= lane)
				fallback = true;		/* a trailing edge without / before its leading edge */
			lpos = eq ? __ffs (eq) - 1 : 0;
		}
		fallback = __any_sync (0xFFFFFFFFu, fallback);
		if (!fallback)
		{
			unsigned v = (s1 ? 1u << lane : 0u) ^ (s0 ? 1u << lpos : 0u);
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const unsigned t = __shfl_up_sync (0xFFFFFFFFu, v, o);
				if (lane >= o) v ^= t;
			}
			/* the active lead with the smallest key (the stack is a few surfaces deep) */
			int best = 0x7FFFFFFF, bl = -1;
			for (unsigned bits = v; bits; bits &= bits - 1u)
			{
				const int j = __ffs (bits) - 1;
				const int kj = sw->lkey[j];
				if (kj < best) { best = kj; bl = j; }
			}
			const int top = bl < 0 ? 1 : (int)(sw->ss[bl] >> 16);
			int pt = __shfl_up_sync (0xFFFFFFFFu, top, 1);
			if (lane == 0) pt = 1;
			const bool chg = lane < n && top != pt;
			const unsigned cm = __ballot_sync (0xFFFFFFFFu, chg);
			/* a change point starts a run that ends at the next change point (the line's end after the last one); the
			 * head run starts at the left edge of the view with the background on top; empty runs are dropped */
			const int iu = uk >> 20;
			const unsigned above = lane == 31 ? 0u : cm & ~((2u << lane) - 1u);
			const int nxt = above ? __ffs (above) - 1 : 0, first = cm ? __ffs (cm) - 1 : 0;
			int en = __shfl_sync (0xFFFFFFFFu, iu, nxt), hen = __shfl_sync (0xFFFFFFFFu, iu, first);
			if (!above) en = tail_u;
			if (!cm) hen = tail_u;
			const bool valid = chg && en > iu;
			const int hvalid = hen > head_u ? 1 : 0;
			const unsigned vm = __ballot_sync (0xFFFFFFFFu, valid);
			nsp = hvalid + __popc (vm);
			__syncwarp ();		/* a.sp shares its memory with the key arrays read above */
			if (valid)
			{
				const int pos = hvalid + __popc (vm & lt);
				sw->a.sp.x0[pos] = (int16_t)iu; sw->a.sp.x1[pos] = (int16_t)en; sw->a.sp.sf[pos] = (uint16_t)top;
			}
			if (lane == 0 && hvalid) { sw->a.sp.x0[0] = (int16_t)head_u; sw->a.sp.x1[0] = (int16_t)hen; sw->a.sp.sf[0] = 1; }
			__syncwarp ();
		}
	}
	else
	{
		/* ---- B2: rank sort (keys are unique) ---- */
		for (int i = lane; i < n; i += 32)
		{
			const unsigned long long ui = sw->a.k.ku[i];
			int rank = 0;
			for (int j = 0; j < n; j++)
			{
				const unsigned long long uj = sw->a.k.ku[j];
				rank += (uj < ui) ? 1 : 0;
				if (uj == ui && j != i)
				{
					/* same u and step: start line and emission order decide (rare) */
					unsigned long long hi_i, lo_i, hi_j, lo_j;
					scan_full_key (sh, sw->a.k.kidx[i], line, &hi_i, &lo_i);
					scan_full_key (sh, sw->a.k.kidx[j], line, &hi_j, &lo_j);
					rank += (lo_j < lo_i) ? 1 : 0;
				}
			}
			sw->su[rank] = (int)((uint32_t)(ui >> 32) ^ 0x80000000u);
			sw->ss[rank] = sh->c_sf[sw->a.k.kidx[i]];
		}
		__syncwarp ();

		/* ---- B3: spans ---- */
		int m = 0;
		for (int base = 0; base < n; base += 32)
		{
			const int k = base + lane;
			int s1 = 0, key = 0, insub = 0;
			if (k < n)
			{
				s1 = (int)(sw->ss[k] >> 16);
				if (s1)
				{
					const int4 g = *reinterpret_cast<const int4 *>(&surfs[s1]);	/* key, flags, face, insubmodel */
					key = g.x; insub = g.w;
				}
			}
			if (insub) fallback = true;
			const unsigned mask = __ballot_sync (0xFFFFFFFFu, s1 != 0);
			const int pos = m + __popc (mask & lt);
			if (s1)
			{
				if (pos < SCAN_LEADS)
				{
					sw->lkey[pos] = key; sw->lsurf[pos] = (uint16_t)s1; sw->lk[pos] = (uint8_t)k;
				}
				sw->lidx[k] = (uint8_t)(pos < SCAN_LEADS ? pos : 0);
			}
			m += __popc (mask);
		}
		if (m > SCAN_LEADS) fallback = true;
		fallback = __any_sync (0xFFFFFFFFu, fallback);
		__syncwarp ();
		if (!fallback)
		{
			/* rank of every leading surface by key */
			for (int i = lane; i < m; i += 32)
			{
				const int ki = sw->lkey[i];
				int r = 0;
				for (int j = 0; j < m; j++)
				{
					const int kj = sw->lkey[j];
					r += (kj < ki) ? 1 : 0;
					if (kj == ki && j != i) fallback = true;	/* a second leading edge of the same surface */
				}
				sw->lrank[i] = (uint8_t)r;
				sw->rsurf[r] = sw->lsurf[i];
				sw->tcount[i] = 0;
			}
			fallback = __any_sync (0xFFFFFFFFu, fallback);
			__syncwarp ();
		}
		if (!fallback)
		{
			/* toggles, XOR scan, top surface after every edge, change points (x0 / sf reuse the key area) */
			unsigned long long carry = 0ull;
			int prevtop = 1, nc = 1;
			if (lane == 0) { sw->a.sp.x0[0] = (int16_t)head_u; sw->a.sp.sf[0] = 1; }
			for (int base = 0; base < n; base += 32)
			{
				const int k = base + lane;
				unsigned long long tog = 0ull;
				int iu = 0;
				if (k < n)
				{
					const uint32_t sfk = sw->ss[k];
					const int s0 = (int)(sfk & 0xFFFF), s1 = (int)(sfk >> 16);
					iu = sw->su[k] >> 20;
					if (s1)
						tog ^= 1ull << sw->lrank[sw->lidx[k]];
					if (s0)
					{
						int found = -1;
						for (int j = 0; j < m; j++)
							if (sw->lsurf[j] == s0) found = j;
						if (found < 0 || (int)sw->lk[found] >= k)
							fallback = true;		/* trailing edge without / before its leading edge */
						else
						{
							tog ^= 1ull << sw->lrank[found];
							atomicAdd (&sw->tcount[found], 1);
						}
					}
				}
				unsigned long long v = tog;
				#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
				{
					const unsigned long long t = __shfl_up_sync (0xFFFFFFFFu, v, o);
					if (lane >= o) v ^= t;
				}
				v ^= carry;
				carry = __shfl_sync (0xFFFFFFFFu, v, 31);
				const int top = v ? (int)sw->rsurf[__ffsll ((long long)v) - 1] : 1;
				int pt = __shfl_up_sync (0xFFFFFFFFu, top, 1);
				if (lane == 0) pt = prevtop;
				prevtop = __shfl_sync (0xFFFFFFFFu, top, 31);
				const bool chg = k < n && top != pt;
				const unsigned mask = __ballot_sync (0xFFFFFFFFu, chg);
				const int pos = nc + __popc (mask & lt);
				if (chg) { sw->a.sp.x0[pos] = (int16_t)iu; sw->a.sp.sf[pos] = (uint16_t)top; }
				nc += __popc (mask);
			}
			__syncwarp ();
			for (int i = lane; i < m; i += 32)
				if (sw->tcount[i] > 1) fallback = true;			/* two trailing edges of one surface */
			fallback = __any_sync (0xFFFFFFFFu, fallback);
			if (!fallback)
			{
				/* runs between change points -> spans, empty runs dropped (in place: writes never pass reads) */
				for (int base = 0; base < nc; base += 32)
				{
					const int j = base + lane;
					int st = 0, en = 0, sf = 0;
					bool valid = false;
					if (j < nc)
					{
						st = sw->a.sp.x0[j]; en = (j + 1 < nc) ? (int)sw->a.sp.x0[j + 1] : tail_u; sf = sw->a.sp.sf[j];
						valid = en > st;
					}
					__syncwarp ();
					const unsigned mask = __ballot_sync (0xFFFFFFFFu, valid);
					const int pos = nsp + __popc (mask & lt);
					if (valid) { sw->a.sp.x0[pos] = (int16_t)st; sw->a.sp.x1[pos] = (int16_t)en; sw->a.sp.sf[pos] = (uint16_t)sf; }
					nsp += __popc (mask);
					__syncwarp ();
				}
			}
		}
	}
	if (fallback)
	{
		/* the reference's own walk, by one lane (r_edge.c:536-563) */
		__syncwarp ();
		if (lane == 0)
		{
			rc_aet_sorted acc; acc.su = sw->su; acc.ss = sw->ss;
			rc_linespans_t ls; ls.surf = sw->a.sp.sf; ls.u0 = sw->a.sp.x0; ls.u1 = sw->a.sp.x1; ls.cap = SPANS;
			nsp = rc_span_walk (&f, vw, surfs, line, acc, n, &ls);
		}
		nsp = __shfl_sync (0xFFFFFFFFu, nsp, 0);
	}
	__syncwarp ();

	/* ---- C: publish ---- */
	if (nsp <= 32)
	{
		/* the common case: one span per lane.  One scan gives the segment offsets and their total, one atomicAdd
		 * reserves the span and segment records; every span lane writes its record and the owners of its segments.
		 * Cell map: the spans that cover at least one whole cell are compacted; a bit per cell marks where such a
		 * span's first whole cell is (one REDUX.OR per 32 cells), so a cell finds its span by a population count
		 * instead of a search. */
		const bool have = lane < nsp;
		const int u0 = have ? (int)sw->a.sp.x0[lane] : 0, x1 = have ? (int)sw->a.sp.x1[lane] : 0, sf = have ? (int)sw->a.sp.sf[lane] : 0;
		const int cnt = x1 - u0, c = have ? rc_span_seg_count (cnt) : 0;
		int incl = c;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const int t = __shfl_up_sync (0xFFFFFFFFu, incl, o);
			if (lane >= o) incl += t;
		}
		const int nseg = __shfl_sync (0xFFFFFFFFu, incl, 31);
		unsigned long long old = 0ull;
		if (lane == 0)
			old = atomicAdd (&f.alloc[0], (unsigned long long)nsp | ((unsigned long long)nseg << 32));
		old = __shfl_sync (0xFFFFFFFFu, old, 0);
		const int spanbase = (int)(old & 0xFFFFFFFFu), segbase = (int)(old >> 32);
		if (spanbase + nsp > f.spancap || segbase + nseg > f.segcap)
		{
			if (lane == 0) atomicOr (f.status, RC_STAT_SPAN_POOL);
			for (int cc = lane; cc < cw; cc += 32)
				cells[cc] = make_int2 (-1, 0);
			return;
		}
		const int myseg = segbase + incl - c;
		const int cs = (u0 + 7) >> 3, ce = x1 >> 3;		/* whole cells cs .. ce-1 */
		const bool full = have && ce > cs;
		const unsigned fm = __ballot_sync (0xFFFFFFFFu, full);
		__syncwarp ();		/* the span arrays were read above: su / ss are reused for the compacted list */
		if (have)
		{
			uint4 *dst = reinterpret_cast<uint4 *>(&f.spans[spanbase + lane]);
			dst[0] = make_uint4 ((uint32_t)(uint16_t)u0 | ((uint32_t)cnt << 16), (uint32_t)((vi * f.height + line) * f.width + u0), (uint32_t)myseg, (uint32_t)(vi * f.surfcap + sf));
			dst[1] = make_uint4 (0u, 0u, 0u, ((uint32_t)line << 20) | (vw->warp_index >= 0 ? RC_SPAN_WARPBIT : 0u));
			surfs[sf].hasspans = 1;
			for (int q = 0; q < c; q++)
				f.segspan[myseg + q] = spanbase + lane;
			if (full)
			{
				const int r = __popc (fm & lt);
				sw->su[r] = myseg;						/* < 2^26 */
				sw->ss[r] = (uint32_t)u0 | ((uint32_t)ce << 11) | ((uint32_t)lane << 21) | ((uint32_t)(x1 & 7) << 26);	/* u0 < 2^11, ce <= 256, lane < 32 */
			}
		}
		__syncwarp ();
		int before = 0;
		for (int base = 0; base < cw; base += 32)
		{
			const unsigned wbits = __reduce_or_sync (0xFFFFFFFFu, (full && (cs >> 5) == (base >> 5)) ? 1u << (cs & 31) : 0u);
			const int cc = base + lane;
			const int k = before + __popc (wbits & (lane == 31 ? 0xFFFFFFFFu : (2u << lane) - 1u)) - 1;
			before += __popc (wbits);
			if (cc < cw)
			{
				int2 cell = make_int2 (-1, 0);
				if (k >= 0)
				{
					const uint32_t pk = sw->ss[k];
					if (cc < (int)((pk >> 11) & 0x3FFu))
					{
						const int pu0 = (int)(pk & 0x7FFu), px1 = (int)(((pk >> 11) & 0x3FFu) << 3) + (int)(pk >> 26);
						cell = make_int2 (rc_cell_si (spanbase + (int)((pk >> 21) & 31u), (cc << 3) - pu0, px1 - pu0),
								  (int)((uint32_t)sw->su[k] * 64u + (uint32_t)((cc << 3) - pu0)));
					}
				}
				cells[cc] = cell;
			}
		}
		return;
	}
	int nseg = 0;
	if (nsp <= SPANS)
		for (int j = lane; j < nsp; j += 32)
			nseg += rc_span_seg_count (sw->a.sp.x1[j] - sw->a.sp.x0[j]);
	#pragma unroll
	for (int o = 16; o; o >>= 1)
		nseg += __shfl_xor_sync (0xFFFFFFFFu, nseg, o);
	unsigned long long old = 0ull;
	if (lane == 0)
		old = atomicAdd (&f.alloc[0], (unsigned long long)nsp | ((unsigned long long)nseg << 32));
	old = __shfl_sync (0xFFFFFFFFu, old, 0);
	const int spanbase = (int)(old & 0xFFFFFFFFu);
	int segbase = (int)(old >> 32);
	if (nsp > SPANS || spanbase + nsp > f.spancap || segbase + nseg > f.segcap)
	{
		if (lane == 0) atomicOr (f.status, RC_STAT_SPAN_POOL);
		for (int c = lane; c < cw; c += 32)
			cells[c] = make_int2 (-1, 0);
		return;
	}
	const int pixrow = (vi * f.height + line) * f.width;
	const uint32_t vflags = ((uint32_t)line << 20) | (vw->warp_index >= 0 ? RC_SPAN_WARPBIT : 0u);
	__syncwarp ();		/* sw->su is reused below */
	for (int base = 0; base < nsp; base += 32)
	{
		const int j = base + lane;
		int u0 = 0, cnt = 0, sf = 0, c = 0;
		if (j < nsp)
		{
			u0 = sw->a.sp.x0[j]; cnt = sw->a.sp.x1[j] - u0; sf = sw->a.sp.sf[j];
			c = rc_span_seg_count (cnt);
		}
		int incl = c;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const int t = __shfl_up_sync (0xFFFFFFFFu, incl, o);
			if (lane >= o) incl += t;
		}
		if (j < nsp)
		{
			uint4 *dst = reinterpret_cast<uint4 *>(&f.spans[spanbase + j]);
			const int myseg = segbase + incl - c;
			dst[0] = make_uint4 ((uint32_t)(uint16_t)u0 | ((uint32_t)cnt << 16), (uint32_t)(pixrow + u0), (uint32_t)myseg, (uint32_t)(vi * f.surfcap + sf));
			dst[1] = make_uint4 (0u, 0u, 0u, vflags);
			sw->su[j] = myseg;
			surfs[sf].hasspans = 1;
		}
		segbase += __shfl_sync (0xFFFFFFFFu, incl, 31);
	}
	__syncwarp ();
	/* owner of every segment of the line: the last span whose first segment is not after it */
	{
		const int seg0 = (int)(old >> 32);
		for (int q = lane; q < nseg; q += 32)
		{
			int lo = 0, hi = nsp - 1, idx = 0;
			while (lo <= hi)
			{
				const int mid = (lo + hi) >> 1;
				if (sw->su[mid] <= seg0 + q) { idx = mid; lo = mid + 1; }
				else hi = mid - 1;
			}
			f.segspan[seg0 + q] = spanbase + idx;
		}
	}
	__syncwarp ();
	/* cell map: the span that covers all of pixels 8c .. 8c+7, if one does (spans are sorted and disjoint) */
	for (int c = lane; c < cw; c += 32)
	{
		const int x = c << 3;
		int lo = 0, hi = nsp - 1, idx = -1;
		while (lo <= hi)
		{
			const int mid = (lo + hi) >> 1;
			if (sw->a.sp.x0[mid] <= x) { idx = mid; lo = mid + 1; }
			else hi = mid - 1;
		}
		int2 cell = make_int2 (-1, 0);
		if (idx >= 0 && x + 8 <= sw->a.sp.x1[idx])
			cell = make_int2 (rc_cell_si (spanbase + idx, x - sw->a.sp.x0[idx], sw->a.sp.x1[idx] - sw->a.sp.x0[idx]),
					  (int)((uint32_t)sw->su[idx] * 64u + (uint32_t)(x - sw->a.sp.x0[idx])));
		cells[c] = cell;
	}
}

/* experiment: one THREAD per (view, scanline), the serial scan of the host simulator (rc_scan_line) */
__global__ void __launch_bounds__(64) k_scan_serial (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	const int line = blockIdx.y * blockDim.x + threadIdx.x;
	if (line < f.height)
		rc_scan_line (&w, &f, blockIdx.x, line);
}

__global__ void k_bmodel (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	RC_GRID_DEP_WAIT ();
	rc_bmodel_face (&w, &f, f.bentbase + blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void k_surfprep (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	RC_GRID_DEP_WAIT ();
	RC_TRACE_SCOPE (3);
	int vi = blockIdx.x;
	int si = rc_surf_slot (&f, vi, blockIdx.y * blockDim.x + threadIdx.x);
	if (si >= 0)
		rc_surf_prep (&w, &f, vi, si);
}

/* Surface-cache build: R_BuildLightmap + R_DrawSurfaceBlock8_mip0..3 (r_light.c:372, r_surf.c:187-375).
 * Work unit = a band of rows of one job holding about RC_CACHE_UNIT_TEXELS texels (the prep kernel lists them: blocks
 * differ 256-fold in size), one WARP per unit, the warps of the grid striding over the list.  What a unit costs is a
 * chain of dependent loads, not arithmetic -- unit -> job -> lightmap samples and source texels -> colormap -> store --
 * so the kernel is built to keep many short chains in flight: everything a unit needs of the face and the texture sits
 * in the job record (rc_surf_prep looked it up), no CTA-wide barrier, no staging of the 16 KB colormap (it is read
 * through L1; staging it per CTA capped the CTAs per SM and cost 16 KB of L2 reads for units of a few KB).
 *   - the rows of blocklights the band needs (17 x 17 luxels at most) are built in the warp's slice of shared memory
 *   - a lane shades 16 consecutive texels of a row: 32-bit loads of source texels, the light of each texel in the
 *     reference's wrapping 32-bit arithmetic in closed form (rc_cache_texel), colormap look-ups, 32-bit stores.
 *     1 B in, 1 B out per texel.  (mip 3 rows, 2-texel blocks, go texel by texel.)
 * History (C2, 6.7 M texels, events): CTA per job with the colormap in shared memory 22 us; CTA per 4096-texel band
 * 27 us (more units, each with a CTA's worth of set-up); the same without the staging, 4 texels per thread 22 us, 16
 * texels per thread 20 us -- a quarter of the instructions and no faster, which is what pointed at the chains. */
#ifndef CACHE_THREADS
#define CACHE_THREADS 256
#endif
#ifndef CACHE_MINB
#define CACHE_MINB 4
#endif
__global__ void __launch_bounds__(CACHE_THREADS, CACHE_MINB) k_cache (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	RC_GRID_DEP_WAIT ();
	RC_TRACE_SCOPE (4);
	__shared__ unsigned s_bl[CACHE_THREADS / 32][18 * 18];
	const uint8_t * __restrict__ cmap = w.colormap;
	const unsigned long long nj = f.alloc[2], nu = f.alloc[8];
	const int njobs = nj > (unsigned long long)f.jobcap ? f.jobcap : (int)nj;
	/* a unit list that overflowed (or jobs that did): whole jobs as units */
	const bool byunit = f.cacheunits != NULL && nu <= (unsigned long long)f.cacheunitcap && nj <= (unsigned long long)f.jobcap;
	const int total = byunit ? (int)nu : njobs;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	unsigned *bl = s_bl[wib];
	const int nwarps = gridDim.x * (CACHE_THREADS / 32);
	for (int un = blockIdx.x * (CACHE_THREADS / 32) + wib; un < total; un += nwarps)
	{
		int j = un, y0 = 0;
		if (byunit)
		{
			const int2 u2 = __ldg (reinterpret_cast<const int2 *>(f.cacheunits) + un);
			j = u2.x; y0 = u2.y;
		}
		const rc_cachejob_t *job = &f.jobs[j];
		const int mip = job->mip, width = job->width, height = job->height, dstofs = job->dstofs;
		const int lw = job->smax, tw = job->tw, th = job->th, soffset = job->soffset, toffset = job->toffset, srcofs = job->srcofs;
		const int rb = job->bandrows;
		const int y1 = byunit ? (y0 + rb < height ? y0 + rb : height) : height;
		const int shift = 4 - mip, bs = 16 >> mip;
		const int lux0 = (y0 >> shift) * lw, lux1 = (((y1 - 1) >> shift) + 2) * lw;	/* luxel rows vb .. vb + 1 of the band's rows */
		const uint8_t *src = w.texels + srcofs;
		uint8_t *dst = f.cachepool + dstofs;
		/* work item: 16 consecutive texels of a row (1, 2 or 4 light blocks at mip 0, 1, 2) */
		const int G = job->items;
		const int nitems = G * (y1 - y0);
		const int qG = job->qG, rG = job->rG;
		const bool words = mip <= 2 && (width & 3) == 0 && (tw & 3) == 0 && (soffset & 3) == 0 && ((((size_t)src) | ((size_t)dst)) & 3) == 0;
		/* source column / row of an item, kept incrementally (no modulo per item): sx = (soffset + 16 cg) % tw, sy = (toffset + y) % th */
		const int dsx = job->dsx, wsx = job->wsx, dsy = job->dsy;
		struct item_t { int row, cg, sx, sy; };
		auto advance = [&] (item_t &t)		/* the item 32 further on */
		{
			t.row += qG; t.cg += rG; t.sx += dsx; t.sy += dsy;
			if (t.cg >= G) { t.cg -= G; t.row++; t.sx -= wsx; if (t.sx < 0) t.sx += tw; if (++t.sy >= th) t.sy -= th; }
			if (t.sx >= tw) t.sx -= tw;
			if (t.sy >= th) t.sy -= th;
		};
		auto fetch = [&] (const item_t &t, bool valid, unsigned (&in)[4])	/* the item's source texels */
		{
			const uint8_t *srow = src + t.sy * tw;
			const int nx = width - (t.cg << 4);
			in[0] = in[1] = in[2] = in[3] = 0;
			if (!(valid && words))
				return;
			if (nx >= 16 && t.sx + 16 <= tw)
			{
				/* the common case: 16 texels that do not wrap */
				#pragma unroll
				for (int wd = 0; wd < 4; wd++)
					in[wd] = __ldg (reinterpret_cast<const unsigned *>(srow + t.sx) + wd);
				return;
			}
			int sxw = t.sx;
			#pragma unroll
			for (int wd = 0; wd < 4; wd++)
				if (wd * 4 < nx)
				{
					while (sxw >= tw) sxw -= tw;
					in[wd] = __ldg (reinterpret_cast<const unsigned *>(srow + sxw));
					sxw += 4;
				}
		};
		auto shade = [&] (const item_t &t, const unsigned (&in)[4])
		{
			const int x0 = t.cg << 4, y = y0 + t.row;
			const int nx = width - x0 < 16 ? width - x0 : 16;
			if (words)
			{
				const int vb = y >> shift, i = y & (bs - 1);
				uint8_t *drow = dst + y * width + x0;
				unsigned light = 0; int lightstep = 0;
				#pragma unroll
				for (int wd = 0; wd < 4; wd++)
				{
					if (wd * 4 < nx)
					{
						const int x = x0 + wd * 4, b = x & (bs - 1);
						if (wd == 0 || b == 0)
						{
							/* a new light block: its row i, stepped right to left (r_surf.c:187-227) */
							const int u = x >> shift;
							const unsigned l00 = bl[vb * lw + u], l01 = bl[vb * lw + u + 1];
							const unsigned l10 = bl[(vb + 1) * lw + u], l11 = bl[(vb + 1) * lw + u + 1];
							const int lightleftstep = (int)((l10 - l00) >> shift);		/* unsigned >>: r_surf.c:202 */
							const int lightrightstep = (int)((l11 - l01) >> shift);
							const unsigned lightleft = l00 + (unsigned)i * (unsigned)lightleftstep;
							const unsigned lightright = l01 + (unsigned)i * (unsigned)lightrightstep;
							lightstep = (int)(lightleft - lightright) >> shift;
							light = lightright + (unsigned)(bs - 1 - b) * (unsigned)lightstep;	/* light of texel x; its right neighbours are one step less each */
						}
						unsigned o = 0;
						#pragma unroll
						for (int k = 0; k < 4; k++)
						{
							o |= (unsigned)__ldg (&cmap[(light & 0xFF00) + ((in[wd] >> (8 * k)) & 0xFF)]) << (8 * k);
							light -= (unsigned)lightstep;
						}
						*reinterpret_cast<unsigned *>(drow + wd * 4) = o;
					}
				}
			}
			else
			{
				for (int k = 0; k < nx; k++)
					dst[y * width + x0 + k] = rc_cache_texel (&w, job, bl, x0 + k, y);
			}
		};
		/* items lane, lane + 32 (a 1024-texel unit has 64): their source texels are requested BEFORE the luxels are built,
		 * so that the two cold reads of a unit -- lightmap samples and texels -- wait side by side */
		item_t ta, tb;
		ta.row = lane / G; ta.cg = lane - ta.row * G;
		ta.sx = (soffset + (ta.cg << 4)) % tw; ta.sy = (toffset + y0 + ta.row) % th;
		tb = ta; advance (tb);
		unsigned ina[4], inb[4];
		fetch (ta, lane < nitems, ina);
		fetch (tb, lane + 32 < nitems, inb);
		__syncwarp ();
		for (int i = lux0 + lane; i < lux1; i += 32)
			bl[i] = rc_cache_luxel (&w, &f, job, i);
		__syncwarp ();
		for (int it = lane; it < nitems; it += 64)
		{
			shade (ta, ina);
			if (it + 32 < nitems)
				shade (tb, inb);
			ta = tb; advance (ta); tb = ta; advance (tb);
			fetch (ta, it + 64 < nitems, ina);
			fetch (tb, it + 96 < nitems, inb);
		}
	}
}

/* Span-segment kernel: one thread per segment of a span (rc_span_seg): eight boundary records, 64 contiguous
 * bytes per thread, 2 KB per warp.  They are staged in shared memory (swizzled) and written out by the whole warp, 512
 * contiguous bytes per store instruction instead of 32 scattered pieces; segments that lie past their span's last
 * subdivision write nothing (rc_span_seg). */
#ifndef SEG_MINB
#define SEG_MINB 8	/* 64 registers: 24.5 us instead of 30 (measured 1 / 6 / 8 / 10 CTAs per SM) */
#endif
__global__ void __launch_bounds__(128, SEG_MINB) k_spanseg (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	RC_GRID_DEP_WAIT ();
	RC_TRACE_SCOPE (5);
	__shared__ uint4 stage[128 * 4];
	unsigned long long a = f.alloc[0];
	int nsegs = (int)(a >> 32);
	if (nsegs > f.segcap) nsegs = f.segcap;
	const int lane = threadIdx.x & 31;
	uint4 *wst = stage + (threadIdx.x & ~31) * 4;
	uint4 *bnd4 = reinterpret_cast<uint4 *>(f.bnd);
	for (int t0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); t0 < nsegs; t0 += gridDim.x * blockDim.x)
	{
		const int t = t0 + lane;
		int wrote = 0;
		if (t < nsegs)
			wrote = rc_span_seg (&w, &f, t, reinterpret_cast<rc_bnd_t *>(wst + lane * 4), lane & 3);
		const unsigned m = __ballot_sync (0xFFFFFFFFu, wrote);
		__syncwarp ();
		#pragma unroll
		for (int r = 0; r < 4; r++)
		{
			const int seg = r * 8 + (lane >> 2), pair = lane & 3;
			if (m & (1u << seg))
				bnd4[(size_t)(t0 + seg) * 4 + pair] = wst[seg * 4 + (pair ^ (seg & 3))];
		}
		__syncwarp ();
	}
}

/* Draw stage.
 * Cell pass: one thread per screen-aligned 8-pixel cell that a single span covers completely.  A warp
 * draws a TILE of 8 cells x 4 rows (64 x 4 pixels): the texel gathers of its 32 lanes then fall into a
 * compact patch of the surface-cache block instead of a 256-pixel strip, which is what bounds this
 * kernel (L1 wavefronts per gather, profiles/).  Per cell the thread needs the cell record
 * (span, subdivision), the span record (1/z, texel block) and two subdivision records (texture
 * coordinates and steps): the cell records of the item after next and the span / subdivision records
 * of the next item are in flight while a cell is drawn, so only the texel gathers are exposed.
 * Common case (cache block or texel blob, no negative 1/z in reach, frame buffer target): straight-line
 * code, one aligned 8-byte colour store and one aligned 16-byte z store.  Everything else (turbulent,
 * sky, solid, warp-buffer views, negative 1/z) goes through the general item rc_draw_pixels.
 * The cells that hold a span boundary are drawn by k_ends. */
#ifndef DRAW_MINB
#define DRAW_MINB 6	/* 80 registers: the prefetched records of the next cell stay in registers (64 spills them) */
#endif
__device__ __noinline__ void draw_general (const rc_world_t *w, const rc_frame_t *f, int si, int rel, bool warpviews)
{
	const rc_span_t sp = rc_load_span (&f->spans[si]);
	rc_draw_pixels (w, f, sp, rel, warpviews);
}

struct draw_item_t {
	int   si;		/* cell record: span | RC_CELL_L0 | RC_CELL_L1, -1: nothing to draw */
	uint32_t sub;		/* cell record: boundary record of the cell's first pixel's subdivision << 3 | position in it */
	uint32_t pix;		/* first pixel of the cell in the colour / z planes */
	uint4 sb;		/* second half of the span record: iziorg, izistep, cacheofs, cwk */
	uint2 b0, b1, b2, b3;	/* boundary records k .. k+3: B[k], B[k+1], then B[k+2] or subdivision k's own step when it is
				 * the span's last, then subdivision k+1's step when that one is */
};

/* subdivisions k and k+1 of a full cell as {s, t, sstep, tstep}, from its four boundary records (rc_sub_of) */
__device__ __forceinline__ void draw_item_subs (const draw_item_t &it, uint4 &r0, uint4 &r1)
{
	const bool l0 = (it.si & RC_CELL_L0) != 0, l1 = ((unsigned)it.si & RC_CELL_L1) != 0;
	r0.x = it.b0.x; r0.y = it.b0.y;
	r0.z = l0 ? it.b2.x : (uint32_t)((int)(it.b1.x - it.b0.x) >> 3);
	r0.w = l0 ? it.b2.y : (uint32_t)((int)(it.b1.y - it.b0.y) >> 3);
	r1.x = it.b1.x; r1.y = it.b1.y;
	r1.z = l1 ? it.b3.x : (uint32_t)((int)(it.b2.x - it.b1.x) >> 3);
	r1.w = l1 ? it.b3.y : (uint32_t)((int)(it.b2.y - it.b1.y) >> 3);
}

#ifndef DRAW_TW
#define DRAW_TW 8		/* tile = DRAW_TW cells x 32 / DRAW_TW rows per warp */
#endif
#ifndef DRAW_PF
#define DRAW_PF 0		/* 0: next item's records prefetched into registers, 1: into L1 (prefetch.global.L1) */
#endif
#define DRAW_TH (32 / DRAW_TW)

__device__ __forceinline__ void draw_prefetch_l1 (const void *p)
{
	asm volatile ("prefetch.global.L1 [%0];" :: "l"(p));
}

/* does the straight-line code handle this full cell?  (cache block or texel blob, frame-buffer target, no negative 1/z in reach) */
__device__ __forceinline__ bool draw_cell_is_fast (const uint4 sb, const uint32_t pix, const bool fast_ok)
{
	const uint32_t izistep = sb.y;
	uint32_t zz = sb.x + pix * izistep, anyneg = (zz - izistep) | zz;
	#pragma unroll
	for (int j = 1; j < 8; j++)
	{
		zz += izistep; anyneg |= zz;
	}
	return fast_ok && (sb.w & 0xE0000u) == 0 && (int)anyneg >= 0;
}

/* the common case of a full cell: D_DrawZSpans + D_DrawSpans8 for 8 screen-aligned pixels */
/* (the eight texels are only gathered here: px[] is packed and stored one tile later by the caller, so that the
 * gathers have a whole iteration to arrive) */
__device__ __forceinline__ bool draw_cell_fast (const rc_world_t &w, const rc_frame_t &f, const uint32_t sub, const uint32_t pix,
	const uint4 sb, const uint4 r0, const uint4 r1, const bool fast_ok, uint32_t (&px)[8])
{
	const unsigned a = sub & 7u;
	const uint32_t izistep = sb.y;
	uint32_t zz[8];
	zz[0] = sb.x + pix * izistep;
	#pragma unroll
	for (int j = 1; j < 8; j++)
		zz[j] = zz[j-1] + izistep;
	const uint32_t anyneg = (zz[0] - izistep) | zz[0] | zz[1] | zz[2] | zz[3] | zz[4] | zz[5] | zz[6] | zz[7];
	const uint32_t cwk = sb.w;
	if (!(fast_ok && (cwk & 0xE0000u) == 0 && (int)anyneg >= 0))
		return false;
	/* D_DrawZSpans: the high halves of consecutive 1/z values, pairwise */
	uint4 zq;
	zq.x = __byte_perm (zz[0], zz[1], 0x7632); zq.y = __byte_perm (zz[2], zz[3], 0x7632);
	zq.z = __byte_perm (zz[4], zz[5], 0x7632); zq.w = __byte_perm (zz[6], zz[7], 0x7632);
	*reinterpret_cast<uint4 *>(f.zb + pix) = zq;
	/* D_DrawSpans8: subdivision k from its pixel a on, subdivision k+1 from cell pixel 8 - a on */
	const uint8_t *pbase = ((cwk & 0x10000u) ? w.texels : f.cachepool) + (int)sb.z;
	const int cachewidth = (int)(cwk & 0xFFFFu);
	uint32_t sstep = r0.z, tstep = r0.w;
	uint32_t ss = r0.x + a * sstep, tt = r0.y + a * tstep;
	#pragma unroll
	for (int j = 0; j < 8; j++)
	{
		if (j > 0 && a + j == 8)
		{
			ss = r1.x; tt = r1.y; sstep = r1.z; tstep = r1.w;
		}
		px[j] = __ldg (&pbase[((int)ss >> 16) + ((int)tt >> 16) * cachewidth]);
		ss += sstep; tt += tstep;
	}
	return true;
}

__device__ __forceinline__ void draw_cell_store (const rc_frame_t &f, const uint32_t pix, const uint32_t (&px)[8])
{
	uint2 q;
	q.x = __byte_perm (__byte_perm (px[0], px[1], 0x0040), __byte_perm (px[2], px[3], 0x0040), 0x5410);
	q.y = __byte_perm (__byte_perm (px[4], px[5], 0x0040), __byte_perm (px[6], px[7], 0x0040), 0x5410);
	*reinterpret_cast<uint2 *>(f.fb + pix) = q;
}

template <bool WARPVIEWS>
__global__ void __launch_bounds__(128, DRAW_MINB) k_draw (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	RC_GRID_DEP_WAIT ();
	RC_TRACE_SCOPE (6);
	const unsigned cw = (unsigned)((f.width + 7) >> 3), tilesx = (cw + DRAW_TW - 1) / DRAW_TW;
	const unsigned row0 = (unsigned)(f.drawview0 * f.height), nrows = (unsigned)((f.drawview1 - f.drawview0) * f.height);
	const unsigned ntiles = tilesx * ((nrows + DRAW_TH - 1) / DRAW_TH);
	const unsigned lane = threadIdx.x & 31, lx = lane % DRAW_TW, ly = lane / DRAW_TW;
	const unsigned wstride = gridDim.x * (blockDim.x >> 5);
	const bool fast_ok = f.zb != NULL && (f.width & 7) == 0;
	const uint4 *spans4 = reinterpret_cast<const uint4 *>(f.spans);
	const uint2 *bnd2 = reinterpret_cast<const uint2 *>(f.bnd);
	const int2 *cells = reinterpret_cast<const int2 *>(f.cellspan);
	/* tile -> (tx, ty), kept incrementally (no divide per item) */
	unsigned tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	unsigned ty = tile / tilesx, tx = tile - ty * tilesx;
	const unsigned dty = wstride / tilesx, dtx = wstride - dty * tilesx;

	/* cell record + first pixel of the tile's cell for this lane; advances (tile, tx, ty) */
	auto fetch_cell = [&] (int2 &cell, uint32_t &pix)
	{
		const unsigned c = tx * DRAW_TW + lx, r = ty * DRAW_TH + ly;
		cell = make_int2 (-1, 0);
		pix = ((row0 + r) * (unsigned)f.width) + (c << 3);
		if (tile < ntiles && c < cw && r < nrows)
			cell = __ldg (&cells[(size_t)(row0 + r) * cw + c]);
		tile += wstride; tx += dtx; ty += dty;
		if (tx >= tilesx) { tx -= tilesx; ty++; }
	};
	auto issue = [&] (draw_item_t &it, const int2 cell, const uint32_t pix)
	{
		it.si = cell.x; it.sub = (uint32_t)cell.y; it.pix = pix;
		const unsigned s = cell.x == -1 ? 0u : (unsigned)RC_CELL_SPAN (cell.x);
		it.sb = __ldg (&spans4[2 * (size_t)s + 1]);
		const unsigned k = it.sub >> 3;
		it.b0 = __ldg (&bnd2[k]); it.b1 = __ldg (&bnd2[k + 1]); it.b2 = __ldg (&bnd2[k + 2]);
		it.b3 = make_uint2 (0u, 0u);
		if ((unsigned)cell.x & RC_CELL_L1)		/* (-1 has the bit too: record 3, harmless) */
			it.b3 = __ldg (&bnd2[k + 3]);
	};

	unsigned t0 = tile;
	unsigned nfull = 0;			/* statistics: cells drawn here or queued */
	uint32_t pxo[8]; uint32_t pendpix = 0; bool pend = false;	/* texels gathered for the previous tile's cell, not stored yet */
	#pragma unroll
	for (int j = 0; j < 8; j++)
		pxo[j] = 0;
	int2 c0, c1; uint32_t p0, p1;
	fetch_cell (c0, p0);
	fetch_cell (c1, p1);
#if DRAW_PF == 0
	draw_item_t cur;
	issue (cur, c0, p0);
#endif
#ifdef DRAW_SKIP_CELLS
	t0 = ntiles;
#endif
	while (t0 < ntiles)
	{
		int2 c2; uint32_t p2;
		fetch_cell (c2, p2);
#if DRAW_PF == 0
		draw_item_t nxt;
		issue (nxt, c1, p1);
#else
		if (c1.x >= 0)
		{
			draw_prefetch_l1 (&spans4[2 * (size_t)c1.x + 1]);
			draw_prefetch_l1 (&bnd2[(uint32_t)c1.y >> 3]);
			draw_prefetch_l1 (&bnd2[((uint32_t)c1.y >> 3) + 3]);
		}
		draw_item_t cur;
		issue (cur, c0, p0);
#endif
		/* cells the straight-line code does not handle are queued for the general kernel (k_ends):
		 * one atomicAdd per warp */
		nfull += cur.si != -1;
		uint32_t pxn[8];
		uint4 r0, r1;
		draw_item_subs (cur, r0, r1);
		const bool fast = cur.si != -1 && draw_cell_fast (w, f, cur.sub, cur.pix, cur.sb, r0, r1, fast_ok, pxn);
		const bool slow = cur.si != -1 && !fast;
		/* colour of the PREVIOUS tile's cell: its gathers were issued one iteration ago */
		if (pend)
			draw_cell_store (f, pendpix, pxo);
		pend = fast; pendpix = cur.pix;
		#pragma unroll
		for (int j = 0; j < 8; j++)
			pxo[j] = pxn[j];
		const unsigned sm = f.slowmode ? 0u : __ballot_sync (0xFFFFFFFFu, slow);
		if (sm)
		{
			unsigned long long base = 0ull;
			if (lane == (unsigned)(__ffs (sm) - 1))
				base = atomicAdd (&f.alloc[3], (unsigned long long)__popc (sm));
			base = __shfl_sync (0xFFFFFFFFu, base, __ffs (sm) - 1);
			if (slow)
				f.slowcells[base + __popc (sm & ((1u << lane) - 1u))] = make_int2 (RC_CELL_SPAN (cur.si), (int)cur.pix);
		}
#if DRAW_PF == 0
		cur = nxt;
#else
		c0 = c1; p0 = p1;
#endif
		c1 = c2; p1 = p2;
		t0 += wstride;
	}
	if (pend)
		draw_cell_store (f, pendpix, pxo);
	#pragma unroll
	for (int o = 16; o; o >>= 1)
		nfull += __shfl_xor_sync (0xFFFFFFFFu, nfull, o);
	if (lane == 0 && nfull)
		atomicAdd (&f.alloc[4], (unsigned long long)nfull);
}

__device__ __noinline__ void draw_mixed_general (const rc_world_t *w, const rc_frame_t *f, int first, int nspans, int cellx, int rowstart, bool warpviews)
{
	rc_draw_mixed_cell (w, f, first, nspans, cellx, rowstart, warpviews);
}

/* Second kernel of the draw stage: the mixed cells (rc_mixed_owner) and the full cells k_draw queued.
 * One item per span; lanes are consecutive spans.  The item loads the records of spans si-1 .. si+2
 * (one 128-byte line) at once: ownership of the cell around the span's first pixel and the list of
 * spans that share that cell (up to three here, else the general code) follow from registers, so the
 * thread has three dependent load levels in all -- span records, subdivision records, texels -- before
 * it stores the cell once.  Shares are merged in registers by the straight-line code of k_draw under a
 * pixel mask; a share that code does not handle sends the whole cell to the general code. */
#ifndef ENDS_MINB
#define ENDS_MINB 4
#endif
template <bool WARPVIEWS>
__global__ void __launch_bounds__(128, ENDS_MINB) k_ends (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	RC_GRID_DEP_WAIT ();
	RC_TRACE_SCOPE (7);
	const uint4 *spans4 = reinterpret_cast<const uint4 *>(f.spans);
	const uint2 *bnd2 = reinterpret_cast<const uint2 *>(f.bnd);
	unsigned long long al = f.alloc[0];
	int nspans = (int)(al & 0xFFFFFFFFu);
	if (nspans > f.spancap) nspans = f.spancap;
#ifdef DRAW_SKIP_ENDS
	nspans = 0;
#endif
	const bool allviews = f.drawview0 == 0 && f.drawview1 == f.nviews;
	for (int si = blockIdx.x * blockDim.x + threadIdx.x; si < nspans; si += gridDim.x * blockDim.x)
	{
		/* level 0: records si-1 .. si+2 */
		uint4 A[4], B[3];
		#pragma unroll
		for (int q = 0; q < 4; q++)
		{
			int idx = si - 1 + q;
			idx = idx < 0 ? 0 : (idx >= nspans ? nspans - 1 : idx);
			A[q] = __ldg (&spans4[2 * (size_t)idx]);
			if (q < 3) B[q] = __ldg (&spans4[2 * (size_t)idx + 1]);
		}
		const int u = (int16_t)(A[1].x & 0xFFFF), count = (int16_t)(A[1].x >> 16), e = u + count;
		if (count <= 0)
			continue;
		const int rowstart = (int)A[1].y - u;
		if (!allviews)
		{
			const int v = (int)A[1].w / f.surfcap;
			if (v < f.drawview0 || v >= f.drawview1)
				continue;
		}
		int uq[4]; bool row[4];
		#pragma unroll
		for (int q = 0; q < 4; q++)
		{
			uq[q] = (int16_t)(A[q].x & 0xFFFF);
			row[q] = (si - 1 + q >= 0) && (si - 1 + q < nspans) && ((int)A[q].y - uq[q] == rowstart) && ((int16_t)(A[q].x >> 16) > 0);
		}
		/* the cell around the span's end (rc_mixed_owner, which = 1): only where the view rectangle ends inside a cell */
		if ((e & 7) && u <= (e & ~7) && !(row[2] && uq[2] < (e & ~7) + 8))
			draw_mixed_general (&w, &f, si, nspans, e & ~7, rowstart, WARPVIEWS);
		/* the cell around its first pixel (which = 0) */
		if (!(u & 7))
			continue;
		const int cellx = u & ~7;
		if (row[0] && uq[0] > cellx)
			continue;			/* an earlier span starts inside the same cell */
		const int firstq = row[0] ? 0 : 1;
		const uint32_t pix = (uint32_t)(rowstart + cellx);
		bool general = f.zb == NULL || (row[3] && uq[3] < cellx + 8);
		/* shares: records firstq .. 2 while they start before the cell's end */
		bool act[3]; int rel[3];
		#pragma unroll
		for (int q = 0; q < 3; q++)
		{
			act[q] = q >= firstq && row[q] && uq[q] < cellx + 8;
			rel[q] = cellx - uq[q];
		}
		/* level 1: subdivisions k, k + 1 of every share from its boundary records (rc_sub_of) */
		uint4 R0[3], R1[3];
		#pragma unroll
		for (int q = 0; q < 3; q++)
		{
			const int k = rel[q] < 0 ? 0 : rel[q] >> 3;
			const int lastsub = (((int16_t)(A[q].x >> 16) + 7) >> 3) - 1;
			const uint2 *rp = bnd2 + (act[q] ? (size_t)A[q].z * 8 + k : 0);
			const uint2 b0 = __ldg (rp), b1 = __ldg (rp + 1), b2 = __ldg (rp + 2), b3 = __ldg (rp + 3);
			const bool l0 = k == lastsub, l1 = k + 1 == lastsub;
			R0[q] = make_uint4 (b0.x, b0.y, l0 ? b2.x : (uint32_t)((int)(b1.x - b0.x) >> 3), l0 ? b2.y : (uint32_t)((int)(b1.y - b0.y) >> 3));
			R1[q] = make_uint4 (b1.x, b1.y, l1 ? b3.x : (uint32_t)((int)(b2.x - b1.x) >> 3), l1 ? b3.y : (uint32_t)((int)(b2.y - b1.y) >> 3));
		}
		uint32_t px[8], zq[8];
		#pragma unroll
		for (int j = 0; j < 8; j++) { px[j] = 0; zq[j] = 0; }
		unsigned mask = 0u;
		#pragma unroll
		for (int q = 0; q < 3; q++)
		{
			if (!act[q] || general)
				continue;
			const int cnt = (int16_t)(A[q].x >> 16);
			const int lo = rel[q] < 0 ? -rel[q] : 0;
			const int hi = cnt - rel[q] < 8 ? cnt - rel[q] : 8;
			if (hi <= lo)
				continue;
			const unsigned vmask = (0xFFu >> (8 - hi)) & (0xFFu << lo);
			const uint32_t izistep = B[q].y;
			uint32_t zz[8];
			zz[0] = B[q].x + pix * izistep;
			#pragma unroll
			for (int j = 1; j < 8; j++)
				zz[j] = zz[j-1] + izistep;
			const uint32_t anyneg = (zz[0] - izistep) | zz[0] | zz[1] | zz[2] | zz[3] | zz[4] | zz[5] | zz[6] | zz[7];
			const uint32_t cwk = B[q].w;
			if ((cwk & 0xE0000u) != 0 || (int)anyneg < 0)
			{
				general = true;
				continue;
			}
			const int a = rel[q] < 0 ? rel[q] : rel[q] & 7;
			const uint8_t *pbase = ((cwk & 0x10000u) ? w.texels : f.cachepool) + (int)B[q].z;
			const int cachewidth = (int)(cwk & 0xFFFFu);
			uint32_t sstep = R0[q].z, tstep = R0[q].w;
			uint32_t ss = R0[q].x + (uint32_t)a * sstep, tt = R0[q].y + (uint32_t)a * tstep;
			#pragma unroll
			for (int j = 0; j < 8; j++)
			{
				if (j > 0 && a + j == 8)
				{
					ss = R1[q].x; tt = R1[q].y; sstep = R1[q].z; tstep = R1[q].w;
				}
				if (vmask & (1u << j))
				{
					px[j] = __ldg (&pbase[((int)ss >> 16) + ((int)tt >> 16) * cachewidth]);
					zq[j] = zz[j];
				}
				ss += sstep; tt += tstep;
			}
			mask |= vmask;
		}
		if (general)
		{
			draw_mixed_general (&w, &f, si - 1 + firstq, nspans, cellx, rowstart, WARPVIEWS);
			continue;
		}
		uint8_t *pd = f.fb + pix;
		int16_t *pz = f.zb + pix;
		if (mask == 0xFFu && (pix & 7u) == 0)
		{
			uint4 z4;
			z4.x = __byte_perm (zq[0], zq[1], 0x7632); z4.y = __byte_perm (zq[2], zq[3], 0x7632);
			z4.z = __byte_perm (zq[4], zq[5], 0x7632); z4.w = __byte_perm (zq[6], zq[7], 0x7632);
			*reinterpret_cast<uint4 *>(pz) = z4;
			uint2 q;
			q.x = __byte_perm (__byte_perm (px[0], px[1], 0x0040), __byte_perm (px[2], px[3], 0x0040), 0x5410);
			q.y = __byte_perm (__byte_perm (px[4], px[5], 0x0040), __byte_perm (px[6], px[7], 0x0040), 0x5410);
			*reinterpret_cast<uint2 *>(pd) = q;
		}
		else
		{
			#pragma unroll
			for (int j = 0; j < 8; j++)
				if (mask & (1u << j))
				{
					pd[j] = (uint8_t)px[j];
					pz[j] = (int16_t)(zq[j] >> 16);
				}
		}
	}
	if (!f.slowmode)
		return;
	/* grouped drawing: k_draw skipped the full cells it does not handle (same test, draw_cell_is_fast) instead of
	 * queueing them; this kernel has issue slots to spare, so it streams over the cell records and draws those */
	{
		const unsigned cw = (unsigned)((f.width + 7) >> 3);
		const unsigned ncells = (unsigned)(f.drawview1 * f.height) * cw;
		const bool fast_ok = f.zb != NULL && (f.width & 7) == 0;
		const int2 *cells = reinterpret_cast<const int2 *>(f.cellspan);
		for (unsigned t = (unsigned)(f.drawview0 * f.height) * cw + blockIdx.x * blockDim.x + threadIdx.x; t < ncells; t += gridDim.x * blockDim.x)
		{
			int2 cell = __ldg (&cells[t]);
			if (cell.x == -1)
				continue;
			cell.x = RC_CELL_SPAN (cell.x);
			const uint4 sb = __ldg (&spans4[2 * (size_t)cell.x + 1]);
			const unsigned row = t / cw, c = t - row * cw;
			const uint32_t pix = row * (unsigned)f.width + (c << 3);
			if (draw_cell_is_fast (sb, pix, fast_ok))
				continue;
			draw_general (&w, &f, cell.x, (int)pix - __ldg (&f.spans[cell.x].pixofs), WARPVIEWS);
		}
	}
}

/* the full cells k_draw left to the general code (turbulent, sky, solid, warp-buffer views, negative 1/z) */
template <bool WARPVIEWS>
__global__ void __launch_bounds__(128) k_slowcells (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	RC_GRID_DEP_WAIT ();
	RC_TRACE_SCOPE (8);
	const unsigned nslow = (unsigned)f.alloc[3];
	for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < nslow; t += gridDim.x * blockDim.x)
	{
		const int2 sc = f.slowcells[t];
		draw_general (&w, &f, sc.x, sc.y - __ldg (&f.spans[sc.x].pixofs), WARPVIEWS);
	}
}

__global__ void k_warp (const __grid_constant__ rc_world_t w, const __grid_constant__ rc_frame_t f)
{
	const int vi = blockIdx.z;
	const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
	rc_warp_pixel (&w, &f, vi, u, v);
}

__global__ void k_alias_verts (const __grid_constant__ rc_frame_t f)
{
	rc_alias_vertex (&f, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

/* The triangles of the batch's alias instances: a thread looks at one (instance, triangle), and the triangles that
 * survive -- about four in ten at C3: half face away, some are clipped off -- are walked (their edges stepped scanline
 * by scanline, the spans going to the queue: rc_pipeline.h "Span queue").
 *   - The slots of the unclipped triangles of a warp are reserved with ONE atomic: a triangle knows its scanline count
 *     before it is walked, and 1.3 M triangles adding to one counter one by one would be the longest thing in the kernel.
 *   - The walk is long, branchy code; with a thread per triangle walking its own, the culled lanes idled through it
 *     (7.4 of 32 lanes active at C3).  So the CTA first COMPACTS its surviving triangles into shared memory and then
 *     walks them with as few, full warps as they need; the index space is (instance, triangle) flattened, so a CTA's
 *     256 candidates do not stop at a model's last triangle. */
#define ATRI_THREADS 256
#ifndef ATRI_CAND
#define ATRI_CAND 1	/* candidates per thread: 2 (the survivors then fill seven of the CTA's eight warps) measured the same, 3 slower */
#endif
__device__ __forceinline__ bool alias_triangle_alive (const rc_frame_t &f, int ii, int ti)
{
	/* what rc_alias_triangle would return from at once: past the model's last triangle, facing away (unclipped:
	 * d_xdenom >= 0, d_polyse.c:150-156), or every vertex beyond the same edge of the view */
	const rc_aliasinst_t *in = &f.alias[ii];
	const rc_aliasmodel_t *m = &f.amodels[in->model];
	if (ti >= m->numtris)
		return false;
	const int *tri = f.atris + 4 * ((size_t)m->tris_ofs + ti);
	const rc_finalvert_t *pfv = f.finalverts + in->vertbase;
	const rc_finalvert_t *a = &pfv[tri[1]], *b = &pfv[tri[2]], *d = &pfv[tri[3]];
	const int clipmask = RC_ALIAS_XY_CLIP_MASK | RC_ALIAS_Z_CLIP;
	if (!in->trivial_accept)
	{
		if (a->flags & b->flags & d->flags & clipmask)
			return false;
		if ((a->flags | b->flags | d->flags) & clipmask)
			return true;		/* clipped: the fan's triangles are tested one by one */
	}
	return (a->v[1] - b->v[1]) * (a->v[0] - d->v[0]) - (a->v[0] - b->v[0]) * (a->v[1] - d->v[1]) < 0;
}

__global__ void __launch_bounds__(ATRI_THREADS) k_alias_tris (const __grid_constant__ rc_frame_t f, const int ninst, const int maxtris)
{
	/* ATRI_CAND candidates per thread: with one, the CTA's survivors fill three or four of its eight warps and the rest
	 * of its register allocation idles through the walk (12 % of the SM's warp slots active at C3) -- which is not what
	 * bounds it: two candidates per thread, seven busy warps, 2.90 against 2.83-2.92 ms alias stage */
	constexpr int NC = ATRI_THREADS * ATRI_CAND;
	__shared__ int s_ii[NC], s_ti[NC], s_rows[NC], s_tri[NC], s_wcount[ATRI_CAND][ATRI_THREADS / 32];
	__shared__ long long s_span0[NC];
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	int ii[ATRI_CAND], ti[ATRI_CAND], rows[ATRI_CAND], tri[ATRI_CAND];
	long long span0[ATRI_CAND];
	bool alive[ATRI_CAND];
	unsigned am[ATRI_CAND];
	#pragma unroll
	for (int c = 0; c < ATRI_CAND; c++)
	{
		const long long gidx = ((long long)blockIdx.x * ATRI_CAND + c) * ATRI_THREADS + threadIdx.x;
		ii[c] = (int)(gidx / maxtris); ti[c] = (int)(gidx - (long long)ii[c] * maxtris);
		const bool valid = ii[c] < ninst;
		rows[c] = valid && f.aspancap > 0 ? rc_alias_triangle_rows (&f, ii[c], ti[c]) : 0;
		alive[c] = valid && (rows[c] > 0 || alias_triangle_alive (f, ii[c], ti[c]));
		/* inclusive prefix sum of the rows, count of the triangles */
		int pre = rows[c];
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const int v = __shfl_up_sync (0xFFFFFFFFu, pre, o);
			if (lane >= (unsigned)o) pre += v;
		}
		const unsigned wants = __ballot_sync (0xFFFFFFFFu, rows[c] > 0);
		const int total = __shfl_sync (0xFFFFFFFFu, pre, 31);
		unsigned long long base = 0ull;
		if (lane == 0 && wants)
			base = atomicAdd (&f.alloc[5], ((unsigned long long)__popc (wants) << RC_AQ_SHIFT) | (unsigned long long)total);
		base = __shfl_sync (0xFFFFFFFFu, base, 0);
		span0[c] = (long long)(base & ((1ull << RC_AQ_SHIFT) - 1)) + (pre - rows[c]);
		tri[c] = (int)(base >> RC_AQ_SHIFT) + __popc (wants & ((1u << lane) - 1u));
		am[c] = __ballot_sync (0xFFFFFFFFu, alive[c]);
		if (lane == 0) s_wcount[c][wid] = __popc (am[c]);
	}
	/* the survivors, in order, into shared memory (binning them by scanline count, so that a warp's triangles are of a
	 * height, was measured and changes nothing: 2.87-2.95 against 2.87-3.00 ms alias stage at C3) */
	__syncthreads ();
	int nitems = 0;
	#pragma unroll
	for (int c = 0; c < ATRI_CAND; c++)
	{
		int off = nitems;
		#pragma unroll
		for (int q = 0; q < ATRI_THREADS / 32; q++)
		{
			if (q < (int)wid) off += s_wcount[c][q];
			nitems += s_wcount[c][q];
		}
		if (alive[c])
		{
			const int k = off + __popc (am[c] & ((1u << lane) - 1u));
			s_ii[k] = ii[c]; s_ti[k] = ti[c]; s_rows[k] = rows[c]; s_tri[k] = tri[c]; s_span0[k] = span0[c];
		}
	}
	__syncthreads ();
	for (int k = threadIdx.x; k < nitems; k += ATRI_THREADS)
		rc_alias_triangle (&f, s_ii[k], s_ti[k], f.aspancap > 0 ? 1 : 0, s_rows[k], s_span0[k], s_tri[k]);
}

/* one thread per queued span: D_PolysetDrawSpans8 */
#ifndef ALIAS_CHUNK
#define ALIAS_CHUNK 4		/* pixels per unit of work of k_alias_spans (a multiple of rc_polyset_span's round of four) */
#endif
#ifndef ALIAS_SPANS_MINB
#define ALIAS_SPANS_MINB 4	/* 64 registers: 3.26 -> 3.08 ms alias stage at C3 against 76 registers (three CTAs per SM); 5 CTAs (48 registers, spills): 3.37 */
#endif
__global__ void __launch_bounds__(256, ALIAS_SPANS_MINB) k_alias_spans (const __grid_constant__ rc_frame_t f)
{
	const unsigned long long a = f.alloc[5];
	long long n = (long long)(a & ((1ull << RC_AQ_SHIFT) - 1));
	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		/* what this batch asked of the queue: the host sizes it for the next one */
		atomicMax (&f.alloc[6], (unsigned long long)n); atomicMax (&f.alloc[7], a >> RC_AQ_SHIFT);
	}
	if (n > f.aspancap) n = f.aspancap;
	const uint4 *sp4 = reinterpret_cast<const uint4 *>(f.aspans);
	long long tested = 0; int front = 0, nsp = 0;
	const uint4 *tr4 = reinterpret_cast<const uint4 *>(f.aspantris);
#ifndef ALIAS_SPAN_PER_THREAD
	/* A thread per span leaves half the lanes idle (15.6 of 32 at C3): a warp's 32 spans run from 1 to 60 pixels and the
	 * warp loops for the longest.  Here the unit of work is a CHUNK of four consecutive pixels (one round of
	 * rc_polyset_span): the warp loads 32 spans, scans their chunk counts, and deals the chunks out one per lane -- the
	 * lane finds its span by a binary search over the scan (shuffles), takes the span's fields from the lane that
	 * loaded it, steps the span's state to the chunk's first pixel in closed form, and runs the reference's stepping
	 * from there.  The closed form is exact: 1/z and light are wrapping 32-bit sums, the skin pointer gains
	 * a_ststepxwhole per pixel plus one carry out of 16 bits of s fraction and one skin row per carry out of 16 bits of
	 * t fraction (d_polyse.c:431-441), and the carries of k steps telescope to floor ((frac + k * step) / 65536) because
	 * both fractions and both steps are below 65536 (k * step < 2^27). */
	const unsigned lane = threadIdx.x & 31u;
	const long long wstride = (long long)gridDim.x * (blockDim.x >> 5) * 32;
	for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < n; base += wstride)
	{
		const long long i = base + lane;
		uint4 s0 = make_uint4 (0u, 0u, 0u, 0u), s1 = make_uint4 (0u, 0u, 0u, 0u);
		if (i < n)
		{
			s0 = __ldcs (&sp4[2 * i]);
			if ((int)s0.y > 0) s1 = __ldcs (&sp4[2 * i + 1]);
		}
		const int mycount = (int)s0.y > 0 ? (int)s0.y : 0;
		const int mychunks = (mycount + ALIAS_CHUNK - 1) / ALIAS_CHUNK;
		int incl = mychunks;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const int v = __shfl_up_sync (0xFFFFFFFFu, incl, o);
			if (lane >= (unsigned)o) incl += v;
		}
		const int total = __shfl_sync (0xFFFFFFFFu, incl, 31);
		const int excl = incl - mychunks;
		tested += mycount; nsp += mycount > 0;
		for (int q0 = 0; q0 < total; q0 += 32)
		{
			const int q = q0 + (int)lane;
			const bool act = q < total;
			const int qq = act ? q : total - 1;
			/* the span whose chunks include qq: the first lane whose inclusive count is above it */
			int lo = 0, hi = 31;
			#pragma unroll
			for (int it = 0; it < 5; it++)
			{
				const int mid = (lo + hi) >> 1;
				const int v = __shfl_sync (0xFFFFFFFFu, incl, mid);
				if (v > qq) hi = mid; else lo = mid + 1;
			}
			const int src = lo;
			const int k0 = (qq - __shfl_sync (0xFFFFFFFFu, excl, src)) * ALIAS_CHUNK;
			rc_aspan_t sp;
			sp.xy = (int)__shfl_sync (0xFFFFFFFFu, s0.x, src); sp.count = (int)__shfl_sync (0xFFFFFFFFu, s0.y, src);
			sp.ptex = (int)__shfl_sync (0xFFFFFFFFu, s0.z, src); sp.stfrac = (int)__shfl_sync (0xFFFFFFFFu, s0.w, src);
			sp.light = (int)__shfl_sync (0xFFFFFFFFu, s1.x, src); sp.zi = (int)__shfl_sync (0xFFFFFFFFu, s1.y, src);
			sp.tri = (int)__shfl_sync (0xFFFFFFFFu, s1.z, src); sp.pad = 0;
			if (!act)
				continue;
			const uint4 t0 = __ldg (&tr4[4 * (size_t)sp.tri]), t1 = __ldg (&tr4[4 * (size_t)sp.tri + 1]),
				    t2 = __ldg (&tr4[4 * (size_t)sp.tri + 2]), t3 = __ldg (&tr4[4 * (size_t)sp.tri + 3]);
			rc_aspantri_t t;
			t.zistepx = (int)t0.x; t.lstepx = (int)t0.y; t.ststepxwhole = (int)t0.z; t.stfracstep = (int)t0.w;
			t.orderbits = (unsigned long long)t1.x | ((unsigned long long)t1.y << 32); t.skinofs = (int)t1.z; t.skinwidth = (int)t1.w;
			t.skinlo = (int)t2.x; t.skinhi = (int)t2.y; t.cmapidx = (int)t2.z; t.viewidx = (int)t2.w;
			t.zresidx = (int)t3.x; t.zwidth = (int)t3.y; t.pad[0] = t.pad[1] = 0;
			/* the span from its pixel k0 on, four pixels at most */
			const int sstep = t.stfracstep & 0xFFFF, tstep = (int)((unsigned)t.stfracstep >> 16);
			const int S = (sp.stfrac & 0xFFFF) + k0 * sstep, T = (int)((unsigned)sp.stfrac >> 16) + k0 * tstep;
			const int left = sp.count - k0;
			sp.xy = (int)(((unsigned)sp.xy & 0xFFFF0000u) | (((unsigned)sp.xy + (unsigned)k0) & 0xFFFFu));
			sp.count = left < ALIAS_CHUNK ? left : ALIAS_CHUNK;
			sp.ptex += k0 * t.ststepxwhole + (S >> 16) + t.skinwidth * (T >> 16);
			sp.stfrac = (S & 0xFFFF) | ((T & 0xFFFF) << 16);
			sp.light = (int)((unsigned)sp.light + (unsigned)k0 * (unsigned)t.lstepx);
			sp.zi = (int)((unsigned)sp.zi + (unsigned)k0 * (unsigned)t.zistepx);
			front += rc_alias_queued_span (&f, sp, t);
		}
	}
#else
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
	{
		const uint4 s0 = __ldcs (&sp4[2 * i]);
		if ((int)s0.y <= 0)
			continue;
		const uint4 s1 = __ldcs (&sp4[2 * i + 1]);
		rc_aspan_t sp;
		sp.xy = (int)s0.x; sp.count = (int)s0.y; sp.ptex = (int)s0.z; sp.stfrac = (int)s0.w;
		sp.light = (int)s1.x; sp.zi = (int)s1.y; sp.tri = (int)s1.z; sp.pad = 0;
		const uint4 t0 = __ldg (&tr4[4 * (size_t)sp.tri]), t1 = __ldg (&tr4[4 * (size_t)sp.tri + 1]),
			    t2 = __ldg (&tr4[4 * (size_t)sp.tri + 2]), t3 = __ldg (&tr4[4 * (size_t)sp.tri + 3]);
		rc_aspantri_t t;
		t.zistepx = (int)t0.x; t.lstepx = (int)t0.y; t.ststepxwhole = (int)t0.z; t.stfracstep = (int)t0.w;
		t.orderbits = (unsigned long long)t1.x | ((unsigned long long)t1.y << 32); t.skinofs = (int)t1.z; t.skinwidth = (int)t1.w;
		t.skinlo = (int)t2.x; t.skinhi = (int)t2.y; t.cmapidx = (int)t2.z; t.viewidx = (int)t2.w;
		t.zresidx = (int)t3.x; t.zwidth = (int)t3.y; t.pad[0] = t.pad[1] = 0;
		front += rc_alias_queued_span (&f, sp, t);
		tested += sp.count; nsp++;
	}
#endif
	if (f.countfrags)
	{
		/* profiling mode: fragments looked at / in front of the world (alloc[9], alloc[10]), spans (alloc[11]) */
		#pragma unroll
		for (int o = 16; o; o >>= 1)
		{
			tested += __shfl_xor_sync (0xFFFFFFFFu, tested, o);
			front += __shfl_xor_sync (0xFFFFFFFFu, front, o);
			nsp += __shfl_xor_sync (0xFFFFFFFFu, nsp, o);
		}
		if ((threadIdx.x & 31) == 0 && nsp)
		{
			atomicAdd (&f.alloc[9], (unsigned long long)tested); atomicAdd (&f.alloc[10], (unsigned long long)front);
			atomicAdd (&f.alloc[11], (unsigned long long)nsp);
		}
	}
}

__global__ void k_sprites (const __grid_constant__ rc_frame_t f)
{
	rc_sprite_span (&f, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void k_particles (const __grid_constant__ rc_frame_t f)
{
	int pi = blockIdx.x * blockDim.x + threadIdx.x;
	if (pi < f.nparticles)
		rc_particle (&f, pi);
}

__global__ void k_zresolve (const __grid_constant__ rc_frame_t f)
{
	rc_zresolve_pixel (&f, blockIdx.z, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y);
}

/* R_EndFrame for batches: 8-bit frame buffers -> 32-bit display pixels, every view through its own
 * palette (what VID_ShiftPalette + the video driver do with the palette R_UpdatePalette built,
 * r_main.c:839-875).  Pure streaming: 1 B read + 4 B written per pixel.  blockIdx.y = view; the view's
 * 1 KB palette sits in shared memory; a thread turns 4 pixels (one 32-bit load) into one 16-byte store, so
 * a warp reads 128 and writes 512 contiguous bytes per instruction; four such groups are in flight per thread. */
__global__ void __launch_bounds__(256) k_present (const uint8_t * __restrict__ fb, uint32_t * __restrict__ out, const uint32_t * __restrict__ pal32, unsigned px)
{
	__shared__ uint32_t spal[256];
	const unsigned vi = blockIdx.y;
	spal[threadIdx.x] = pal32[vi * 256 + threadIdx.x];
	__syncthreads ();
	const unsigned stride = gridDim.x * blockDim.x;
	unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
	if ((px & 3u) || (reinterpret_cast<size_t>(fb) & 3) || (reinterpret_cast<size_t>(out) & 15))
	{
		/* a plane size that is not a multiple of 4 pixels puts every view after the first off the 4 / 16-byte
		 * alignment the vector path needs (as does an unaligned slice from the caller): one pixel per thread */
		for (; q < px; q += stride)
			out[(size_t)vi * px + q] = spal[fb[(size_t)vi * px + q]];
		return;
	}
	const uint32_t *src = reinterpret_cast<const uint32_t *>(fb + (size_t)vi * px);
	uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)vi * px);
	const unsigned nq = px >> 2;
	for (; q + 3 * stride < nq; q += 4 * stride)
	{
		uint32_t v[4];
		#pragma unroll
		for (int i = 0; i < 4; i++)
			v[i] = __ldcs (&src[q + i * stride]);
		#pragma unroll
		for (int i = 0; i < 4; i++)
			__stcs (&dst[q + i * stride], make_uint4 (spal[v[i] & 0xFF], spal[(v[i] >> 8) & 0xFF], spal[(v[i] >> 16) & 0xFF], spal[v[i] >> 24]));
	}
	for (; q < nq; q += stride)
	{
		const uint32_t v = __ldcs (&src[q]);
		__stcs (&dst[q], make_uint4 (spal[v & 0xFF], spal[(v >> 8) & 0xFF], spal[(v >> 16) & 0xFF], spal[v >> 24]));
	}
}

/* 2D overlay (r_draw.c): one CTA per view runs the view's commands IN ORDER (later commands overwrite earlier ones),
 * its threads share the pixels of each command. */
__global__ void __launch_bounds__(256) k_draw2d (const rc_d2ctx_t x, const rc_draw2d_t * __restrict__ cmds, const int * __restrict__ first, uint8_t *fb)
{
	const int vi = blockIdx.x;
	uint8_t *plane = fb + (size_t)vi * x.width * x.height;
	for (int ci = first[vi]; ci < first[vi + 1]; ci++)
	{
		const rc_draw2d_t c = cmds[ci];
		const int cnt = rc_d2_count (&x, &c);
		for (int p = threadIdx.x; p < cnt; p += blockDim.x)
			rc_d2_pixel (&x, &c, p, plane);
		__syncthreads ();
	}
}

__global__ void k_fill64 (unsigned long long *p, unsigned long long v, size_t n)
{
	RC_GRID_DEP_WAIT ();
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		p[i] = v;
}

__global__ void k_flush_cache (unsigned long long *keys, size_t n, unsigned long long *poolptr, unsigned long long first)
{
	RC_GRID_DEP_WAIT ();
	RC_TRACE_SCOPE (9);
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		keys[i] = RC_CACHE_EMPTY;
	if (blockIdx.x == 0 && threadIdx.x == 0)
		*poolptr = first;
}

/* start of a chunk: the span / job / queued-cell / full-cell counters (alloc[0], alloc[2..4]; alloc[1] is the surface
 * cache's) and, for the first chunk of a batch, the status words -- one launch instead of three memsets */
__global__ void k_reset_chunk (unsigned long long *alloc, int *status)
{
	RC_GRID_DEP_WAIT ();
	const unsigned t = threadIdx.x;
#ifdef RC_TRACE
	g_trace[t] = (t & 1) ? 0ull : ~0ull;
	__syncwarp ();
	{ RC_TRACE_SCOPE (0); }
#endif
	if (t == 0 || (t >= 2 && t <= 5) || t == 8)
		alloc[t] = 0ull;
	if (status && t < 16)
		status[t] = 0;
	if (status && (t == 6 || t == 7 || (t >= 9 && t <= 11)))
		alloc[t] = 0ull;		/* first chunk of a batch: the span queue's high-water marks */
}

/* status word of a pipelined call, stored straight into pinned host memory: a device->host cudaMemcpyAsync would
 * queue on the copy engine behind the previous batch's frame-buffer copies and, being in stream order, hold the
 * next batch's kernels back until those copies are through */
__global__ void k_status_out (const int *d_status, int *host_status)
{
	RC_GRID_DEP_WAIT ();
	*host_status = *d_status;
	__threadfence_system ();
}

/* ------------------------------------------------------------------------- host ------- */
template <typename T>
static int upload (rc_gpu_t *g, const T *src, size_t count, const T **dst, char *err, size_t errlen)
{
	void *d = NULL;
	size_t bytes = sizeof(T) * (count ? count : 1);
	CK (cudaMalloc (&d, bytes));
	g->worldallocs.push_back (d);
	if (count) CK (cudaMemcpy (d, src, sizeof(T) * count, cudaMemcpyHostToDevice));
	*dst = (const T *)d;
	return RC_OK;
}

extern "C" {

rc_gpu_t *rc_gpu_create (int device, int width, int height, size_t cache_bytes, int max_views, char *err, size_t errlen)
{
	int count = 0;
	cudaError_t e = cudaGetDeviceCount (&count);
	if (e != cudaSuccess || count <= 0)
	{
		snprintf (err, errlen, "ref_cuda needs a CUDA device: %s (there is no CPU rendering path)",
			  e != cudaSuccess ? cudaGetErrorString (e) : "no device found");
		return NULL;
	}
	if (device < 0 || device >= count) { snprintf (err, errlen, "device %d out of range (%d devices)", device, count); return NULL; }
	if (cudaSetDevice (device) != cudaSuccess) { snprintf (err, errlen, "cudaSetDevice(%d) failed", device); return NULL; }
	rc_gpu_t *g = new rc_gpu_s ();
	memset (&g->stats, 0, sizeof(g->stats));
	g->device = device; g->width = width; g->height = height;
	g->haveworld = false; g->skybox = 0; g->chunkcap = 0; g->maxchunk = 0; g->maxviews = max_views; g->spans_per_line = 24;
	for (int i = 0; i < RC_ASYNC_DEPTH; i++)
	{
		g->d_afb[i] = NULL; g->d_azb[i] = NULL; g->aoutcap[i] = 0;
		cudaEventCreateWithFlags (&g->adone_s[i], cudaEventDisableTiming);
		cudaEventCreateWithFlags (&g->adone_copy[i], cudaEventDisableTiming);
	}
	g->d_d2blob = NULL; g->d_d2pics = NULL; g->d_d2cmds = NULL; g->d_d2first = NULL; g->d2cmdcap = 0; g->d2firstcap = 0;
	g->palvalid = 0;
	g->d_pal32 = NULL; g->palcap = 0; g->h_pal32[0] = g->h_pal32[1] = NULL; g->palslot = 0;
	cudaEventCreateWithFlags (&g->evpal[0], cudaEventDisableTiming); cudaEventCreateWithFlags (&g->evpal[1], cudaEventDisableTiming);
	g->d_presentin = NULL; g->d_presentout = NULL; g->presentcap = 0;
	g->astatus = NULL;
	if (cudaHostAlloc (&g->astatus, sizeof(int) * RC_ASYNC_DEPTH, cudaHostAllocDefault) != cudaSuccess || !g->astatus)
	{
		snprintf (err, errlen, "cudaHostAlloc (status words) failed: %s", cudaGetErrorString (cudaGetLastError ()));
		delete g;
		return NULL;
	}
	g->ahead = 0; g->ainflight = 0; g->amode = 0;
	{ const char *e = getenv ("RC_ASYNC_GROUPS"); g->agroups = e && atoi (e) >= 1 && atoi (e) <= 8 ? atoi (e) : 4; }
	g->d_views = NULL; g->d_fb = NULL; g->d_zb = NULL; g->outcap = 0; g->profiling = 0; g->last_n = 0; g->flush_pending = 0; g->launches_total = 0; g->d_dlights = NULL; g->dlightcap = 0; g->d_warpbuf = NULL; g->warpcap = 0;
	g->d_amodels = NULL; g->d_stverts = g->d_atris = NULL; g->d_poseverts = g->d_skinpixels = NULL; g->d_anormdots = NULL;
	g->maxverts = g->maxtris = 0; g->d_alias = NULL; g->aliascap = 0; g->d_finalverts = NULL; g->finalvertcap = 0;
	g->d_particles = NULL; g->particlecap = 0; g->d_sprites = NULL; g->spritecap = 0; g->d_spritepixels = NULL; g->d_colormaps = NULL; g->colormapcap = 0; g->d_zres = NULL; g->zrescap = 0; g->d_aspans = NULL; g->aspancap = 0; g->d_aspantris = NULL; g->aspantricap = 0; g->aspan_need = 0; g->aspantri_need = 0; g->d_zscratch = NULL; g->zscratchcap = 0; g->cur_batch = NULL;
	g->d_bents = NULL; g->bentcap = 0; g->d_leafkey = NULL; g->d_nbsurfs = NULL; g->bsurfcap = 0;
	cudaFuncSetAttribute (k_scan<SCAN_SMALL_AET, SCAN_SMALL_CAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(scan_cta_t<SCAN_SMALL_AET, SCAN_SMALL_CAND>));
	cudaFuncSetAttribute (k_scan<256, SCAN_CAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(scan_cta_t<256, SCAN_CAND>));
	cudaFuncSetAttribute (k_front<FRONT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
	cudaFuncSetAttribute (k_front<FRONT_MANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
	cudaDeviceProp prop;
	cudaGetDeviceProperties (&prop, device);
	g->numsms = prop.multiProcessorCount;
	cudaStreamCreateWithFlags (&g->stream, cudaStreamNonBlocking);
	cudaStreamCreateWithFlags (&g->copystream, cudaStreamNonBlocking);
	for (int i = 0; i < 8; i++) cudaEventCreateWithFlags (&g->evgroup[i], cudaEventDisableTiming);
	cudaEventCreateWithFlags (&g->evcopy, cudaEventDisableTiming);
	cudaStreamCreateWithFlags (&g->sidestream, cudaStreamNonBlocking);
	cudaEventCreateWithFlags (&g->evfork, cudaEventDisableTiming);
	cudaEventCreateWithFlags (&g->evjoin, cudaEventDisableTiming);
	for (int i = 0; i < 2; i++) { g->h_viewstage[i] = NULL; g->viewstagecap[i] = 0; g->viewstage_used[i] = 0; cudaEventCreateWithFlags (&g->ev_viewstage[i], cudaEventDisableTiming); }
	g->viewstage_idx = 0;
	g->mirrorgroups = 0; g->mirror_deferred = 0;
	for (int i = 0; i < 4; i++) { g->mpend[i].fb = NULL; g->mpend[i].valid = 0; cudaEventCreateWithFlags (&g->mpend[i].ev, cudaEventDisableTiming); }
	cudaEventCreateWithFlags (&g->evrender, cudaEventDisableTiming);
	/* Environment knobs, all read ONCE here; they select measured-and-rejected variants for A/B runs and are not part of
	 * the API: RC_ASYNC_GROUPS=1..8 groups of a pipelined call (default 4), RC_SLOW_IN_ENDS k_ends finds the
	 * general cells itself, RC_SCAN_LARGE start with the large scan instance, RC_NO_GRAPH no CUDA-graph replay,
	 * RC_FRONT_512 512-thread front-end CTAs on many-view batches */
	g->front_512 = getenv ("RC_FRONT_512") ? 1 : 0;
	g->front_256 = getenv ("RC_FRONT_256") ? 1 : 0;
	g->use_pdl = getenv ("RC_NO_PDL") ? 0 : 1;
	{ const char *e = getenv ("RC_FRONT_CSIZE"); g->front_csize = e ? atoi (e) : 0; if (g->front_csize != 1 && g->front_csize != 2 && g->front_csize != 4 && g->front_csize != 8) g->front_csize = 0; }
	g->scan_serial = getenv ("RC_SCAN_SERIAL") ? 1 : 0;
	g->slow_in_ends = getenv ("RC_SLOW_IN_ENDS") ? 1 : 0;	/* measured: 14 us slower than the queue + k_slowcells at C2 */
	g->scan_large = getenv ("RC_SCAN_LARGE") ? 1 : 0;
	g->draw_first = getenv ("RC_DRAW_FIRST") ? 1 : 0;
	{ const char *e = getenv ("RC_DRAW_CTAS"); g->draw_ctas = e && atoi (e) >= 1 && atoi (e) <= 32 ? atoi (e) : 16; }
	{ const char *e = getenv ("RC_ENDS_CTAS"); g->ends_ctas = e && atoi (e) >= 1 && atoi (e) <= 32 ? atoi (e) : 16; }
	for (int i = 0; i < RC_GRAPHS; i++) { g->graphs[i].exec = NULL; g->graphs[i].used = 0; }
	g->graph_clock = 0; g->use_graphs = getenv ("RC_NO_GRAPH") ? 0 : 1;
	g->host_fb = NULL; g->host_z = NULL; g->mirror_fb = NULL; g->mirror_z = NULL; g->groupfn = NULL; g->groupuser = NULL; g->groupstream = NULL; g->ngroups = 0;
	for (int i = 0; i < 16; i++) cudaEventCreate (&g->ev[i]);

	std::vector<int> st (RC_SIN_BUFFER_SIZE), ist (RC_SIN_BUFFER_SIZE);
	RC_BuildTurbTables (st.data (), ist.data (), RC_SIN_BUFFER_SIZE);
	cudaMalloc (&g->d_sintable, sizeof(int) * RC_SIN_BUFFER_SIZE);
	cudaMalloc (&g->d_intsintable, sizeof(int) * RC_SIN_BUFFER_SIZE);
	cudaMemcpy (g->d_sintable, st.data (), sizeof(int) * RC_SIN_BUFFER_SIZE, cudaMemcpyHostToDevice);
	cudaMemcpy (g->d_intsintable, ist.data (), sizeof(int) * RC_SIN_BUFFER_SIZE, cudaMemcpyHostToDevice);

	g->cachecap = cache_bytes ? (long long)cache_bytes : (1ll << 30);
	if (g->cachecap > 0x7FFF0000ll) g->cachecap = 0x7FFF0000ll;	/* offsets are int */
	g->cacheslots = 1 << 18;
	if (cudaMalloc (&g->d_cachepool, g->cachecap) != cudaSuccess ||
	    cudaMalloc (&g->d_cachekeys, sizeof(unsigned long long) * g->cacheslots) != cudaSuccess ||
	    cudaMalloc (&g->d_cachevals, sizeof(int) * g->cacheslots) != cudaSuccess ||
	    cudaMalloc (&g->d_alloc, sizeof(unsigned long long) * 16) != cudaSuccess ||
	    cudaMalloc (&g->d_status, sizeof(int) * 16) != cudaSuccess)
	{
		snprintf (err, errlen, "cudaMalloc (surface cache) failed: %s", cudaGetErrorString (cudaGetLastError ()));
		delete g;
		return NULL;
	}
	cudaMemset (g->d_alloc, 0, sizeof(unsigned long long) * 16);
	cudaMemset (g->d_cachekeys, 0xFF, sizeof(unsigned long long) * g->cacheslots);
	cudaDeviceSynchronize ();
	return g;
}

static void free_chunk (rc_gpu_t *g)
{
	cudaFree (g->d_views); cudaFree (g->d_facelist); cudaFree (g->d_nfaces); cudaFree (g->d_facepos);
	cudaFree (g->d_facemark); cudaFree (g->d_nodevisit); cudaFree (g->d_viewleaf);
	cudaFree (g->d_nbsurfs); cudaFree (g->d_leafkey); g->d_nbsurfs = NULL; g->d_leafkey = NULL;
	cudaFree (g->d_edges); cudaFree (g->d_nedges); cudaFree (g->d_surfs); cudaFree (g->d_spans);
	cudaFree (g->d_cellspan); cudaFree (g->d_slowcells); cudaFree (g->d_jobs); cudaFree (g->d_cacheunits); g->d_cacheunits = NULL; cudaFree (g->d_segspan); cudaFree (g->d_bnd);
	g->d_views = NULL; g->chunkcap = 0;
}

static void free_world (rc_gpu_t *g)
{
	for (void *p : g->worldallocs) cudaFree (p);
	g->worldallocs.clear ();
	g->haveworld = false;
}

void rc_gpu_destroy (rc_gpu_t *g)
{
	if (!g) return;
	cudaSetDevice (g->device);
	cudaDeviceSynchronize ();
	free_chunk (g); free_world (g);
	cudaFree (g->d_cachepool); cudaFree (g->d_cachekeys); cudaFree (g->d_cachevals);
	cudaFree (g->d_alloc); cudaFree (g->d_status); cudaFree (g->d_sintable); cudaFree (g->d_intsintable);
	cudaFree (g->d_fb); cudaFree (g->d_zb); cudaFree (g->d_dlights); cudaFree (g->d_warpbuf);
	for (int i = 0; i < RC_ASYNC_DEPTH; i++) { cudaFree (g->d_afb[i]); cudaFree (g->d_azb[i]); cudaEventDestroy (g->adone_s[i]); cudaEventDestroy (g->adone_copy[i]); }
	cudaFreeHost (g->astatus);
	cudaFree (g->d_pal32); cudaFree (g->d_presentin); cudaFree (g->d_presentout);
	cudaFree (g->d_d2blob); cudaFree (g->d_d2pics); cudaFree (g->d_d2cmds); cudaFree (g->d_d2first);
	cudaFreeHost (g->h_pal32[0]); cudaFreeHost (g->h_pal32[1]); cudaEventDestroy (g->evpal[0]); cudaEventDestroy (g->evpal[1]);
	cudaFree (g->d_sprites); cudaFree (g->d_spritepixels);
	cudaFree (g->d_bents); cudaFree (g->d_alias); cudaFree (g->d_finalverts); cudaFree (g->d_particles);
	cudaFree (g->d_colormaps); cudaFree (g->d_zres); cudaFree (g->d_zscratch); cudaFree (g->d_aspans); cudaFree (g->d_aspantris);
	cudaFree (g->d_amodels); cudaFree (g->d_stverts); cudaFree (g->d_atris); cudaFree (g->d_poseverts);
	cudaFree (g->d_skinpixels); cudaFree (g->d_anormdots);
	for (int i = 0; i < RC_GRAPHS; i++) if (g->graphs[i].exec) cudaGraphExecDestroy (g->graphs[i].exec);
	for (int i = 0; i < 16; i++) cudaEventDestroy (g->ev[i]);
	for (int i = 0; i < 8; i++) cudaEventDestroy (g->evgroup[i]);
	cudaEventDestroy (g->evcopy); cudaEventDestroy (g->evfork); cudaEventDestroy (g->evjoin);
	for (int i = 0; i < 2; i++) { cudaEventDestroy (g->ev_viewstage[i]); if (g->h_viewstage[i]) cudaFreeHost (g->h_viewstage[i]); }
	for (int i = 0; i < 4; i++) cudaEventDestroy (g->mpend[i].ev);
	cudaEventDestroy (g->evrender);
	cudaStreamDestroy (g->copystream); cudaStreamDestroy (g->sidestream);
	cudaStreamDestroy (g->stream);
	delete g;
}

int rc_gpu_upload_alias (rc_gpu_t *g, rc_hostalias_t **models, int n, const float *anorm_dots, char *err, size_t errlen)
{
	std::vector<rc_aliasmodel_t> am;
	std::vector<int> stv, tri;
	std::vector<uint8_t> pose, skin;
	cudaSetDevice (g->device);
	cudaDeviceSynchronize ();
	g->maxverts = g->maxtris = 0;
	for (int i = 0; i < n; i++)
	{
		rc_hostalias_t *a = models[i];
		rc_aliasmodel_t m;
		m.numverts = a->numverts; m.numtris = a->numtris; m.skinwidth = a->skinwidth; m.skinheight = a->skinheight;
		m.stverts_ofs = (int)stv.size () / 4; m.tris_ofs = (int)tri.size () / 4;
		m.poses_ofs = (int)pose.size () / 4; m.skins_ofs = (int)skin.size ();
		stv.insert (stv.end (), a->stverts, a->stverts + 4 * a->numverts);
		tri.insert (tri.end (), a->tris, a->tris + 4 * a->numtris);
		pose.insert (pose.end (), a->poseverts, a->poseverts + 4 * (size_t)a->numposes * a->numverts);
		m.skins_len = a->skinimagelen; m.pad[0] = m.pad[1] = m.pad[2] = 0;
		skin.insert (skin.end (), a->skinpixels, a->skinpixels + (size_t)a->skinimagelen);
		am.push_back (m);
		a->device_index = i;
		if (a->numverts > g->maxverts) g->maxverts = a->numverts;
		if (a->numtris > g->maxtris) g->maxtris = a->numtris;
	}
	cudaFree (g->d_amodels); cudaFree (g->d_stverts); cudaFree (g->d_atris); cudaFree (g->d_poseverts);
	cudaFree (g->d_skinpixels); cudaFree (g->d_anormdots);
	g->d_amodels = NULL; g->d_stverts = g->d_atris = NULL; g->d_poseverts = g->d_skinpixels = NULL; g->d_anormdots = NULL;
	CK (cudaMalloc (&g->d_amodels, sizeof(rc_aliasmodel_t) * (am.size () + 1)));
	CK (cudaMalloc (&g->d_stverts, sizeof(int) * (stv.size () + 4)));
	CK (cudaMalloc (&g->d_atris, sizeof(int) * (tri.size () + 4)));
	CK (cudaMalloc (&g->d_poseverts, pose.size () + 16));
	CK (cudaMalloc (&g->d_skinpixels, skin.size () + 16));
	CK (cudaMalloc (&g->d_anormdots, sizeof(float) * 16 * 256));
	if (n)
	{
		CK (cudaMemcpy (g->d_amodels, am.data (), sizeof(rc_aliasmodel_t) * am.size (), cudaMemcpyHostToDevice));
		CK (cudaMemcpy (g->d_stverts, stv.data (), sizeof(int) * stv.size (), cudaMemcpyHostToDevice));
		CK (cudaMemcpy (g->d_atris, tri.data (), sizeof(int) * tri.size (), cudaMemcpyHostToDevice));
		CK (cudaMemcpy (g->d_poseverts, pose.data (), pose.size (), cudaMemcpyHostToDevice));
		CK (cudaMemcpy (g->d_skinpixels, skin.data (), skin.size (), cudaMemcpyHostToDevice));
	}
	CK (cudaMemcpy (g->d_anormdots, anorm_dots, sizeof(float) * 16 * 256, cudaMemcpyHostToDevice));
	return RC_OK;
}

int rc_gpu_upload_sprites (rc_gpu_t *g, rc_hostsprite_t **models, int n, char *err, size_t errlen)
{
	std::vector<uint8_t> pix;
	cudaSetDevice (g->device);
	cudaDeviceSynchronize ();
	for (int i = 0; i < n; i++)
	{
		models[i]->device_ofs = (int)pix.size ();
		pix.insert (pix.end (), models[i]->pixels, models[i]->pixels + models[i]->pixelbytes);
	}
	cudaFree (g->d_spritepixels); g->d_spritepixels = NULL;
	CK (cudaMalloc (&g->d_spritepixels, pix.size () + 16));
	if (!pix.empty ())
		CK (cudaMemcpy (g->d_spritepixels, pix.data (), pix.size (), cudaMemcpyHostToDevice));
	return RC_OK;
}

int rc_gpu_upload_world (rc_gpu_t *g, const rc_hostworld_t *hw, int maxedges, int maxsurfs, char *err, size_t errlen)
{
	const rc_world_t *h = &hw->w;
	cudaSetDevice (g->device);
	cudaDeviceSynchronize ();
	free_world (g);
	free_chunk (g);
	g->w = *h;
#define UP(field, count) if (upload (g, h->field, (size_t)(count), &g->w.field, err, errlen) != RC_OK) return RC_ERR_CUDA;
	UP (planes, h->numplanes + 6)
	UP (nodes, h->numnodes)
	UP (leafs, h->numleafs)
	UP (faces, h->numfaces)
	UP (surfedges, h->numsurfedges + 24)
	UP (separtner, h->numsurfedges + 1)
	UP (fedges, h->numsurfedges + 24 + 1)
	UP (edges, (size_t)(h->numedges + 13) * 2)
	UP (vertexes, (size_t)(h->numvertexes + 8) * 3)
	UP (texinfo, h->numtexinfo + 6)
	UP (textures, h->numtextures + 1)
	UP (texels, hw->texelbytes + 16)
	UP (lightdata, hw->lightbytes + 16)
	UP (marksurfaces, h->nummarksurfaces + 1)
	UP (markleaf, h->nummarksurfaces + 1)
	UP (nodevis, (size_t)h->visrowwords * h->numleafs)
	UP (colormap, 256 * 64)
	UP (nodeaux, h->numnodes + 1)
	UP (leafdfs, h->numleafs + 1)
#undef UP
	g->w.sintable = g->d_sintable;
	g->w.intsintable = g->d_intsintable;
	g->w.drawskybox = g->skybox;
	g->maxedges = maxedges; g->maxsurfs = maxsurfs;
	g->haveworld = true;
	rc_gpu_flush_caches (g);
	return RC_OK;
}

void rc_gpu_flush_caches (rc_gpu_t *g)
{
	g->flush_pending = 1;
}

int rc_gpu_device_alloc (rc_gpu_t *g, size_t bytes, void **ptr)
{
	cudaSetDevice (g->device);
	return cudaMalloc (ptr, bytes) == cudaSuccess ? RC_OK : RC_ERR_CUDA;
}

void rc_gpu_device_free (rc_gpu_t *g, void *ptr) { cudaSetDevice (g->device); cudaFree (ptr); }

int rc_gpu_ipc_export (rc_gpu_t *g, void *ptr, unsigned char *handle)
{
	cudaIpcMemHandle_t h;
	static_assert (sizeof(h) == 64, "CUDA IPC handle size");
	cudaSetDevice (g->device);
	if (cudaIpcGetMemHandle (&h, ptr) != cudaSuccess) return RC_ERR_CUDA;
	memcpy (handle, &h, sizeof(h));
	return RC_OK;
}

int rc_gpu_ipc_import (rc_gpu_t *g, const unsigned char *handle, void **ptr)
{
	cudaIpcMemHandle_t h;
	cudaSetDevice (g->device);
	memcpy (&h, handle, sizeof(h));
	return cudaIpcOpenMemHandle (ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? RC_OK : RC_ERR_CUDA;
}

void rc_gpu_ipc_close (rc_gpu_t *g, void *ptr) { cudaSetDevice (g->device); cudaIpcCloseMemHandle (ptr); }
void rc_gpu_set_mirror_groups (rc_gpu_t *g, int ngroups) { g->mirrorgroups = ngroups; }
void rc_gpu_set_mirror (rc_gpu_t *g, void *fb, void *z) { g->mirror_fb = (uint8_t *)fb; g->mirror_z = (int16_t *)z; }
void rc_gpu_set_mirror_deferred (rc_gpu_t *g, int on) { g->mirror_deferred = on != 0; }
int rc_gpu_mirror_fence (rc_gpu_t *g, void *stream, char *err, size_t errlen)
{
	cudaSetDevice (g->device);
	cudaStream_t s = stream ? (cudaStream_t)stream : cudaStreamLegacy;
	for (int i = 0; i < 4; i++)
		if (g->mpend[i].valid)
			CK (cudaStreamWaitEvent (s, g->mpend[i].ev, 0));
	return RC_OK;
}

void rc_gpu_set_group_callback (rc_gpu_t *g, rc_group_fn fn, void *user, void *side_stream, int ngroups)
{
	g->groupfn = fn; g->groupuser = user; g->groupstream = (cudaStream_t)side_stream; g->ngroups = ngroups;
}

int rc_gpu_set_skybox (rc_gpu_t *g, const uint8_t *pixels, char *err, size_t errlen)
{
	cudaSetDevice (g->device);
	cudaDeviceSynchronize ();
	if (pixels)
		CK (cudaMemcpy (g->d_cachepool, pixels, RC_SKYBOX_BYTES, cudaMemcpyHostToDevice));
	g->skybox = pixels != NULL;
	g->w.drawskybox = g->skybox;
	return RC_OK;
}

static void do_flush (rc_gpu_t *g, cudaStream_t s)
{
	/* only the slots that can be in use are cleared: the table is sized for the pool */
	/* the pool starts after the sky-box pictures (the allocation pointer is reset by the same kernel: one item in
	 * stream order instead of a kernel and a copy) */
	k_flush_cache<<<g->numsms * 4, 256, 0, s>>> (g->d_cachekeys, (size_t)g->cacheslots, g->d_alloc + 1, (unsigned long long)RC_SKYBOX_BYTES);
	g->flush_pending = 0;
}

static size_t per_view_bytes (const rc_gpu_t *g, int facecap, int edgecap, int surfcap, int numfaces)
{
	size_t lines = g->height;
	size_t cw = (g->width + 7) / 8;
	size_t spans = lines * g->spans_per_line, segs = spans + spans / 4 + lines * (cw / 8 + 1);
	return sizeof(rc_view_t) + sizeof(rc_facerec_t) * facecap + 6 * (size_t)numfaces + 8 * (size_t)g->w.numnodes + sizeof(rc_edge_t) * edgecap +
	       sizeof(rc_surf_t) * surfcap + sizeof(rc_span_t) * spans + (4 + 8 * sizeof(rc_bnd_t)) * segs + (sizeof(rc_cell_t) + 8) * cw * lines + sizeof(rc_cachejob_t) * 8 + 16;
}

static int ensure_chunk (rc_gpu_t *g, int n, int want_bsurfs, char *err, size_t errlen)
{
	const rc_world_t *w = &g->w;
	int facecap = (w->numworldfaces < 65000 ? w->numworldfaces : 65000) + 6;	/* + the six sky-box faces */
	/* brush-entity surf_t slots, once allocated, are kept: batches that alternate between having brush entities and not
	 * must not free and reallocate every chunk buffer (and re-instantiate the graph) each time */
	if (g->d_views && g->bsurfcap > want_bsurfs) want_bsurfs = g->bsurfcap;
	int edgecap = g->maxedges, surfcap = facecap + 2 + want_bsurfs;
	size_t freeb = 0, totalb = 0;
	if (g->d_views && g->facecap == facecap && g->edgecap == edgecap && g->bsurfcap == want_bsurfs && (g->chunkcap >= n || g->chunkcap == g->maxchunk))
		return RC_OK;
	cudaMemGetInfo (&freeb, &totalb);
	size_t pv = per_view_bytes (g, facecap, edgecap, surfcap, w->numfaces);
	size_t budget = (size_t)(totalb * 0.35);
	int maxchunk = (int)(budget / pv);
	if (maxchunk < 1) maxchunk = 1;
	if (maxchunk > 65535) maxchunk = 65535;		/* one grid dimension per view */
	{
		/* pixel indices inside a chunk are 32-bit (rc_span_t.pixofs) */
		long long pxcap = 0x7FFFFFFFll / ((long long)g->width * g->height);
		if (maxchunk > pxcap) maxchunk = (int)(pxcap > 0 ? pxcap : 1);
	}
	if (g->maxviews > 0 && maxchunk > g->maxviews) maxchunk = g->maxviews;
	int want = n < maxchunk ? n : maxchunk;
	g->maxchunk = maxchunk;
	free_chunk (g);
	g->facecap = facecap; g->edgecap = edgecap; g->surfcap = surfcap; g->bsurfcap = want_bsurfs;
	size_t lines = (size_t)g->height * want;
	g->spancap = (int)(lines * g->spans_per_line < 0x7FFFFFFF ? lines * g->spans_per_line : 0x7FFFFFFF);
	const size_t cw = (g->width + 7) / 8;
	size_t segs = (size_t)g->spancap + (size_t)g->spancap / 4 + lines * (cw / 8 + 1);	/* a span needs ((nsub + 2) + 7) / 8 segments (rc_span_seg_count) */
	g->segcap = (int)(segs < (1 << 26) - 2 ? segs : (1 << 26) - 2);	/* rc_cell_t.sub holds segment * 64 + pixel in 32 bits */
	g->jobcap = want * 8 + 4096;
	if (g->jobcap > want * surfcap) g->jobcap = want * surfcap;
	CK (cudaMalloc (&g->d_views, sizeof(rc_view_t) * want));
	CK (cudaMalloc (&g->d_facelist, sizeof(rc_facerec_t) * (size_t)want * facecap));
	CK (cudaMalloc (&g->d_nfaces, sizeof(int) * want));
	CK (cudaMalloc (&g->d_facepos, sizeof(uint16_t) * (size_t)want * w->numfaces));
	CK (cudaMalloc (&g->d_facemark, sizeof(int) * (size_t)want * w->numfaces));
	CK (cudaMalloc (&g->d_nodevisit, sizeof(rc_nodevisit_t) * (size_t)want * w->numnodes));
	CK (cudaMalloc (&g->d_viewleaf, sizeof(int) * want));
	CK (cudaMalloc (&g->d_nbsurfs, sizeof(int) * want));
	CK (cudaMalloc (&g->d_leafkey, sizeof(int) * (size_t)want * w->numleafs));
	CK (cudaMalloc (&g->d_edges, sizeof(rc_edge_t) * (size_t)want * edgecap));
	CK (cudaMalloc (&g->d_nedges, sizeof(int) * want));
	CK (cudaMalloc (&g->d_surfs, sizeof(rc_surf_t) * (size_t)want * surfcap));
	CK (cudaMalloc (&g->d_spans, sizeof(rc_span_t) * (size_t)g->spancap));
	CK (cudaMalloc (&g->d_cellspan, sizeof(rc_cell_t) * lines * cw));
	CK (cudaMalloc (&g->d_slowcells, sizeof(int2) * lines * cw));
	CK (cudaMalloc (&g->d_segspan, sizeof(int) * (size_t)g->segcap));
	CK (cudaMalloc (&g->d_bnd, sizeof(rc_bnd_t) * (8 * (size_t)g->segcap + 16)));
	CK (cudaMalloc (&g->d_jobs, sizeof(rc_cachejob_t) * (size_t)g->jobcap));
	g->cacheunitcap = g->jobcap * 8 + 65536;
	CK (cudaMalloc (&g->d_cacheunits, sizeof(int2) * (size_t)g->cacheunitcap));
	g->chunkcap = want;
	return RC_OK;
}

/* one chunk of views, everything asynchronous on `s` */
static int run_chunk (rc_gpu_t *g, const rc_view_t *hviews, int first_view, int n, uint8_t *fb, int16_t *zb, cudaStream_t s,
		      bool timed, char *err, size_t errlen)
{
	bool anywarp = false;
	for (int i = 0; i < n; i++) anywarp |= hviews[i].warp_index >= 0;
	const rc_world_t &w = g->w;
	rc_frame_t f;
	memset (&f, 0, sizeof(f));
	f.nviews = n; f.width = g->width; f.height = g->height; f.views = g->d_views; f.dlights = g->d_dlights;
	f.viewbase = first_view;
	f.bents = g->d_bents; f.nbents = g->cur_batch ? g->cur_batch->nbents : 0;
	f.leafkey = g->d_leafkey; f.nbsurfs = g->d_nbsurfs; f.bsurfcap = g->bsurfcap;
	f.facelist = g->d_facelist; f.facecap = g->facecap; f.nfaces = g->d_nfaces; f.facepos = g->d_facepos;
	f.facemark = g->d_facemark; f.nodevisit = g->d_nodevisit; f.viewleaf = g->d_viewleaf;
	f.edges = g->d_edges; f.edgecap = g->edgecap; f.nedges = g->d_nedges;
	f.surfs = g->d_surfs; f.surfcap = g->surfcap; f.maxedges = g->maxedges; f.maxsurfs = g->maxsurfs;
	f.spans = g->d_spans; f.spancap = g->spancap; f.cellspan = g->d_cellspan; f.slowcells = g->d_slowcells;
	f.segspan = g->d_segspan; f.bnd = g->d_bnd; f.segcap = g->segcap;
	f.alloc = g->d_alloc;
	f.cachekeys = g->d_cachekeys; f.cachevals = g->d_cachevals; f.cachemask = g->cacheslots - 1;
	f.cachepool = g->d_cachepool; f.cachecap = g->cachecap;
	f.jobs = g->d_jobs; f.jobcap = g->jobcap; f.cacheunits = reinterpret_cast<int *>(g->d_cacheunits); f.cacheunitcap = g->cacheunitcap;
	f.fb = fb; f.zb = zb; f.status = g->d_status; f.warpbuf = g->d_warpbuf;
	f.countfrags = g->profiling ? 1 : 0;

	/* The plain device path (no entities, no warp views, no per-group consumers, no profiling) is a fixed sequence of
	 * memsets and kernels on two streams whose parameters only change with the world or the buffers: it is captured
	 * into a CUDA graph once and relaunched, which removes the launch gaps and the event round trips of the forks. */
	const bool capturable = g->use_graphs && !(g->rflags & RC_RF_DEFER) && !timed && s != cudaStreamLegacy && !g->host_fb && !g->groupfn && !anywarp &&
		!(g->cur_batch && (g->cur_batch->nzres || g->cur_batch->nbents));
	rc_frame_t f0;
	memcpy (&f0, &f, sizeof(f));
	struct capture_guard_t {
		cudaStream_t s; bool on;
		~capture_guard_t () { if (on) { cudaGraph_t gr = NULL; cudaStreamEndCapture (s, &gr); if (gr) cudaGraphDestroy (gr); } }
	} capture = { s, false };
	int gslot = 0;
	if (timed) cudaEventRecord (g->ev[0], s);
	if (g->rflags & RC_RF_DEFER)
	{
		/* a copy from pageable memory holds the calling thread until the stream has reached it, i.e. until the previous
		 * part is drawn (C4, eight parts: 22.7 of the call's 29.7 ms were spent inside these calls) */
		const int b = (g->viewstage_idx ^= 1);
		const size_t bytes = sizeof(rc_view_t) * (size_t)n;
		if (g->viewstage_used[b]) CK (cudaEventSynchronize (g->ev_viewstage[b]));
		if (g->viewstagecap[b] < bytes)
		{
			if (g->h_viewstage[b]) cudaFreeHost (g->h_viewstage[b]);
			g->h_viewstage[b] = NULL; g->viewstagecap[b] = 0;
			CK (cudaHostAlloc ((void **)&g->h_viewstage[b], bytes, cudaHostAllocDefault));
			g->viewstagecap[b] = bytes;
		}
		memcpy (g->h_viewstage[b], hviews, bytes);
		CK (cudaMemcpyAsync (g->d_views, g->h_viewstage[b], bytes, cudaMemcpyHostToDevice, s));
		CK (cudaEventRecord (g->ev_viewstage[b], s));
		g->viewstage_used[b] = 1;
	}
	else
		CK (cudaMemcpyAsync (g->d_views, hviews, sizeof(rc_view_t) * n, cudaMemcpyHostToDevice, s));
	if (capturable)
	{
		gslot = 0;
		for (int i = 0; i < RC_GRAPHS; i++)
		{
			if (g->graphs[i].exec && !memcmp (&g->graphs[i].w, &w, sizeof(w)) && !memcmp (&g->graphs[i].f, &f0, sizeof(f0)))
			{
				g->graphs[i].used = ++g->graph_clock;
				CK (cudaGraphLaunch (g->graphs[i].exec, s));
				g->stats.kernel_launches += 9;	/* the eight pipeline kernels + k_reset_chunk */
				return RC_OK;
			}
			if (g->graphs[i].used < g->graphs[gslot].used) gslot = i;	/* least recently used (an empty slot has 0) */
		}
		if (g->graphs[gslot].exec) { cudaGraphExecDestroy (g->graphs[gslot].exec); g->graphs[gslot].exec = NULL; g->graphs[gslot].used = 0; }
		CK (cudaStreamBeginCapture (s, cudaStreamCaptureModeThreadLocal));
		capture.on = true;
	}
	k_reset_chunk<<<1, 32, 0, s>>> (g->d_alloc, (first_view == 0 && !(g->rflags & RC_RF_KEEP_STATUS)) ? g->d_status : NULL);

	{
		/* cluster size: as many CTAs per view as fit the machine in one wave, at most 8 */
		int csize = 1;
		while (csize < 8 && n * csize * 2 <= g->numsms * FRONT_CTAS_PER_SM) csize *= 2;	/* one wave */
		cudaLaunchConfig_t cfg;
		cudaLaunchAttribute attr[1];
		memset (&cfg, 0, sizeof(cfg));
		bool manyviews = n > g->numsms && !g->front_512;	/* more views than SMs: several waves */
		if (g->front_csize > 0) csize = g->front_csize;		/* experiments: RC_FRONT_CSIZE / RC_FRONT_256 */
		if (g->front_256) manyviews = true;
		cfg.gridDim = dim3 (n * csize); cfg.blockDim = dim3 (manyviews ? FRONT_MANY : FRONT_THREADS); cfg.stream = s;
		{
			const size_t fs = manyviews ? FRONT_SMEM_OF (FRONT_MANY) : FRONT_SMEM_OF (FRONT_THREADS);
			cfg.dynamicSmemBytes = fs > 2 * (size_t)w.numnodes + 16 ? fs : 2 * (size_t)w.numnodes + 16;
		}
		attr[0].id = cudaLaunchAttributeClusterDimension;
		attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
		cfg.attrs = attr; cfg.numAttrs = 1;
		cudaLaunchAttribute attr2[2];
		attr2[0] = attr[0];
		attr2[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr2[1].val.programmaticStreamSerializationAllowed = 1;
		if (g->use_pdl && !timed) { cfg.attrs = attr2; cfg.numAttrs = 2; }
		if (manyviews) CK (cudaLaunchKernelEx (&cfg, k_front<FRONT_MANY>, w, f));
		else CK (cudaLaunchKernelEx (&cfg, k_front<FRONT_THREADS>, w, f));
	}
	if (timed) cudaEventRecord (g->ev[1], s);
	if (g->cur_batch && g->cur_batch->nbents)
	{
		const rc_batch_t *b = g->cur_batch;
		int blo = 0, bhi;
		while (blo < b->nbents && b->bents[blo].view < first_view) blo++;
		bhi = blo;
		while (bhi < b->nbents && b->bents[bhi].view < first_view + n) bhi++;
		for (int first = blo; first < bhi; first += 32768)
		{
			int cnt = bhi - first < 32768 ? bhi - first : 32768;
			rc_frame_t f2 = f;
			f2.bentbase = first;	/* surf_t.entity indexes f.bents in the prep kernel: same base everywhere */
			dim3 gb ((b->maxbfaces + 31) / 32, cnt);
			k_bmodel<<<gb, 32, 0, s>>> (w, f2);
			g->stats.kernel_launches++;
		}
	}
	if (timed) cudaEventRecord (g->ev[2], s);
	dim3 gs (n, (g->height + SCAN_WARPS - 1) / SCAN_WARPS);
	if (g->scan_serial) { dim3 gss (n, (g->height + 63) / 64); k_scan_serial<<<gss, 64, 0, s>>> (w, f); }
	else if (g->scan_large) k_scan<256, SCAN_CAND><<<gs, SCAN_WARPS * 32, sizeof(scan_cta_t<256, SCAN_CAND>), s>>> (w, f);
	else CK (launch_k (k_scan<SCAN_SMALL_AET, SCAN_SMALL_CAND>, gs, dim3 (SCAN_WARPS * 32), sizeof(scan_cta_t<SCAN_SMALL_AET, SCAN_SMALL_CAND>), s, g->use_pdl && !timed, w, f));
	if (timed) cudaEventRecord (g->ev[3], s);
	dim3 gp (n, (g->surfcap + 63) / 64);
	CK (launch_k (k_surfprep, gp, dim3 (64), 0, s, g->use_pdl && !timed, w, f));
	if (timed) cudaEventRecord (g->ev[4], s);
	if (timed)
	{
		k_cache<<<g->numsms * CACHE_MINB, CACHE_THREADS, 0, s>>> (w, f);
		cudaEventRecord (g->ev[5], s);
		k_spanseg<<<g->numsms * 8, 128, 0, s>>> (w, f);
		cudaEventRecord (g->ev[7], s);
	}
	else
	{
		/* the surface-cache build and the span-segment kernel only depend on the prep kernel: they run
		 * side by side (second stream forked and joined with events)
		 * (k_cache first: its one wave of CTAs thins out as the warps run out of units and k_spanseg's CTAs move in, 6 us
		 * after it on the device's clock; the other way round -- the longer kernel on the main stream, programmatically
		 * dependent -- k_spanseg's eight CTAs per SM keep k_cache out until they drain and the pair ends at the same
		 * time: RC_TRACE build, tools/variant.sh) */
		cudaEventRecord (g->evfork, s);
		cudaStreamWaitEvent (g->sidestream, g->evfork, 0);
		k_spanseg<<<g->numsms * 8, 128, 0, g->sidestream>>> (w, f);
		cudaEventRecord (g->evjoin, g->sidestream);
		CK (launch_k (k_cache, dim3 (g->numsms * CACHE_MINB), dim3 (CACHE_THREADS), 0, s, g->use_pdl, w, f));
		cudaStreamWaitEvent (s, g->evjoin, 0);
	}
	const bool entities = g->cur_batch && g->cur_batch->nzres;
	bool grouped = false;
	if ((g->host_fb || g->groupfn) && !entities && !anywarp && n >= 8)
	{
		/* the views are drawn in groups and every finished group is handed to its consumer on a second
		 * stream while the next one is drawn: the D2H copy of the host-output path (PCIe: the longest single
		 * item of a batch), or whatever the caller's callback enqueues (bench.py: the NCCL gather) */
		grouped = true;
		/* host output crosses PCIe (the copy of a group is longer than its draw: four groups); a mirror in peer memory
		 * is an NVLink copy, short next to the per-launch cost of a group: two */
		const int mgroups = g->mirrorgroups >= 1 && g->mirrorgroups <= 8 ? (g->mirrorgroups <= n ? g->mirrorgroups : n) : 2;
		const int ngroups = g->host_fb ? (g->host_fb == g->mirror_fb ? mgroups : (g->amode ? g->agroups : (n >= 32 ? 4 : 2))) : (g->ngroups > 0 && g->ngroups <= 8 ? (g->ngroups <= n ? g->ngroups : n) : 4);
		cudaStream_t side = g->host_fb ? g->copystream : g->groupstream;
		const size_t px = (size_t)g->width * g->height;
		/* the mixed cells of ALL views in one launch on the second stream, beside the first groups' full cells
		 * (latency bound, about as long as the draw of two groups): the first group waits for it */
		f.drawview0 = 0; f.drawview1 = n; f.slowmode = 1;
		cudaEventRecord (g->evfork, s);
		cudaStreamWaitEvent (g->sidestream, g->evfork, 0);
		k_ends<false><<<g->numsms * 16, 128, 0, g->sidestream>>> (w, f);
		cudaEventRecord (g->evjoin, g->sidestream);
		for (int gi = 0; gi < ngroups; gi++)
		{
			f.drawview0 = (int)((long long)n * gi / ngroups); f.drawview1 = (int)((long long)n * (gi + 1) / ngroups);
			const int nv = f.drawview1 - f.drawview0;
			int ctas = (int)(((size_t)nv * g->height * ((g->width + 7) / 8) + 127) / 128);
			if (ctas > g->numsms * 16) ctas = g->numsms * 16;
			CK (launch_k (k_draw<false>, dim3 (ctas), dim3 (128), 0, s, g->use_pdl && !timed && gi > 0, w, f));	/* (group 0 follows a fork) */
			if (timed) cudaEventRecord (g->ev[8], s);
			if (gi == 0) cudaStreamWaitEvent (s, g->evjoin, 0);
			cudaEventRecord (g->evgroup[gi], s);
			cudaStreamWaitEvent (side, g->evgroup[gi], 0);
			if (g->host_fb)
			{
				const size_t ofs = px * (size_t)(first_view + f.drawview0);
				CK (cudaMemcpyAsync (g->host_fb + ofs, fb + px * f.drawview0, px * nv, cudaMemcpyDefault, side));
				if (g->host_z)
					CK (cudaMemcpyAsync (g->host_z + ofs, zb + px * f.drawview0, px * nv * sizeof(int16_t), cudaMemcpyDefault, side));
			}
			else
				g->groupfn (g->groupuser, first_view + f.drawview0, nv);
		}
		g->stats.kernel_launches += ngroups - 2;	/* ngroups draws + one k_ends, no k_slowcells; 8 are counted below */
		f.slowmode = 0;
		if (g->host_fb && g->amode)
			cudaEventRecord (g->adone_copy[g->amode - 1], side);	/* the caller waits for it (rc_gpu_wait_host); the next call's kernels do not */
		else if (g->host_fb)
		{
			cudaEventRecord (g->evcopy, side);
			cudaStreamWaitEvent (s, g->evcopy, 0);		/* the stream's work is complete only when the copies are */
		}
	}
	else
	{
		f.drawview0 = 0; f.drawview1 = n;
		if (timed)
		{
			if (anywarp) k_draw<true><<<g->numsms * 16, 128, 0, s>>> (w, f);
			else k_draw<false><<<g->numsms * 16, 128, 0, s>>> (w, f);
			cudaEventRecord (g->ev[8], s);
			if (anywarp) { k_ends<true><<<g->numsms * 16, 128, 0, s>>> (w, f); k_slowcells<true><<<g->numsms * 8, 128, 0, s>>> (w, f); }
			else { k_ends<false><<<g->numsms * 16, 128, 0, s>>> (w, f); k_slowcells<false><<<g->numsms * 8, 128, 0, s>>> (w, f); }
		}
		else
		{
			/* mixed cells on the second stream, beside the full cells (disjoint pixels; both only read the
			 * span / subdivision records and the cache) */
			/* the cells k_draw leaves to the general code are queued for k_slowcells; letting k_ends find them itself
			 * (slowmode 1, as grouped drawing does to save launches) measured slower here */
			f.slowmode = g->slow_in_ends;
			cudaEventRecord (g->evfork, s);
			cudaStreamWaitEvent (g->sidestream, g->evfork, 0);
			if (g->draw_first)
			{
				if (anywarp) k_draw<true><<<g->numsms * g->draw_ctas, 128, 0, s>>> (w, f);
				else k_draw<false><<<g->numsms * g->draw_ctas, 128, 0, s>>> (w, f);
			}
			if (anywarp) k_ends<true><<<g->numsms * g->ends_ctas, 128, 0, g->sidestream>>> (w, f);
			else k_ends<false><<<g->numsms * g->ends_ctas, 128, 0, g->sidestream>>> (w, f);
			cudaEventRecord (g->evjoin, g->sidestream);
			if (!g->draw_first)
			{
				if (anywarp) k_draw<true><<<g->numsms * g->draw_ctas, 128, 0, s>>> (w, f);
				else k_draw<false><<<g->numsms * g->draw_ctas, 128, 0, s>>> (w, f);
			}
			if (!f.slowmode)
			{
				if (anywarp) CK (launch_k (k_slowcells<true>, dim3 (g->numsms * 8), dim3 (128), 0, s, g->use_pdl, w, f));
				else CK (launch_k (k_slowcells<false>, dim3 (g->numsms * 8), dim3 (128), 0, s, g->use_pdl, w, f));
			}
			cudaStreamWaitEvent (s, g->evjoin, 0);
			f.slowmode = 0;
		}
	}
	if (timed) cudaEventRecord (g->ev[10], s);		/* end of the world's draw stage */
	if (g->cur_batch && g->cur_batch->nzres)
	{
		const rc_batch_t *b = g->cur_batch;
		f.alias = g->d_alias; f.nalias = b->nalias; f.amodels = g->d_amodels; f.stverts = g->d_stverts; f.atris = g->d_atris;
		f.poseverts = g->d_poseverts; f.skinpixels = g->d_skinpixels; f.anormdots = g->d_anormdots;
		f.colormaps = g->d_colormaps; f.finalverts = g->d_finalverts; f.particles = g->d_particles; f.nparticles = b->nparticles;
		f.zres = g->d_zres;
		/* instances / particles of this chunk's views (both arrays are ordered by view) */
		int alo = 0, ahi = b->nalias, plo = 0, phi = b->nparticles;
		while (alo < b->nalias && b->alias[alo].view < first_view) alo++;
		ahi = alo;
		while (ahi < b->nalias && b->alias[ahi].view < first_view + n) ahi++;
		while (plo < b->nparticles && b->particles[plo].view < first_view) plo++;
		phi = plo;
		while (phi < b->nparticles && b->particles[phi].view < first_view + n) phi++;
		/* instances per launch: grid.y (at most 65535), and the 28 bits the queue's counter has for triangle records */
		int astep = 32768;
		while (astep > 1 && (long long)astep * (g->maxtris > 0 ? g->maxtris : 1) * 8 > (1ll << 28)) astep >>= 1;
		for (int first = alo; first < ahi; first += astep)
		{
			int cnt = ahi - first < astep ? ahi - first : astep;
			rc_frame_t f2 = f;
			f2.alias = g->d_alias + first;
			dim3 ga ((g->maxverts + 127) / 128, cnt);
			const unsigned gt = (unsigned)(((long long)cnt * (g->maxtris > 0 ? g->maxtris : 1) + ATRI_THREADS * ATRI_CAND - 1) / (ATRI_THREADS * ATRI_CAND));
			f2.aspans = g->d_aspans; f2.aspancap = g->aspancap; f2.aspantris = g->d_aspantris; f2.aspantricap = g->aspantricap;
			CK (cudaMemsetAsync (g->d_alloc + 5, 0, sizeof(unsigned long long), s));
			k_alias_verts<<<ga, 128, 0, s>>> (f2);
			k_alias_tris<<<gt, ATRI_THREADS, 0, s>>> (f2, cnt, g->maxtris > 0 ? g->maxtris : 1);
			k_alias_spans<<<g->numsms * 8, 256, 0, s>>> (f2);
			g->stats.kernel_launches += 3;
		}
		{
			/* sprite entities of this chunk's views (ordered by view) */
			int slo = 0, shi;
			while (slo < b->nsprites && b->sprites[slo].view < first_view) slo++;
			shi = slo;
			while (shi < b->nsprites && b->sprites[shi].view < first_view + n) shi++;
			for (int first = slo; first < shi; first += 32768)
			{
				int cnt = shi - first < 32768 ? shi - first : 32768;
				rc_frame_t f4 = f;
				f4.sprites = g->d_sprites + first; f4.nsprites = cnt; f4.spritepixels = g->d_spritepixels;
				dim3 gsprite ((g->height + 1 + 63) / 64, cnt);
				k_sprites<<<gsprite, 64, 0, s>>> (f4);
				g->stats.kernel_launches++;
			}
		}
		if (timed) cudaEventRecord (g->ev[11], s);	/* alias models + sprites */
		if (phi > plo)
		{
			rc_frame_t f3 = f;
			f3.particles = g->d_particles + plo; f3.nparticles = phi - plo;
			k_particles<<<(phi - plo + 127) / 128, 128, 0, s>>> (f3);
			g->stats.kernel_launches++;
		}
		if (timed) cudaEventRecord (g->ev[12], s);	/* particles */
		dim3 gz ((g->width + 127) / 128, g->height, n);
		k_zresolve<<<gz, 128, 0, s>>> (f);
		g->stats.kernel_launches++;
		if (timed) cudaEventRecord (g->ev[13], s);	/* z resolve */
	}
	if (anywarp)
	{
		dim3 gw2 ((g->width + 127) / 128, g->height, n);
		k_warp<<<gw2, 128, 0, s>>> (w, f);
		g->stats.kernel_launches++;
	}
	if (g->groupfn && !g->host_fb && !grouped)
	{
		/* one group: the whole chunk, after the entity / warp kernels */
		cudaEventRecord (g->evgroup[0], s);
		cudaStreamWaitEvent (g->groupstream, g->evgroup[0], 0);
		g->groupfn (g->groupuser, first_view, n);
	}
	if (g->host_fb && !grouped)
	{
		/* entity / warp kernels touched the frame after the draw: one copy at the end of the chunk */
		const size_t px = (size_t)g->width * g->height;
		CK (cudaMemcpyAsync (g->host_fb + px * first_view, fb, px * n, cudaMemcpyDefault, s));
		if (g->host_z)
			CK (cudaMemcpyAsync (g->host_z + px * first_view, zb, px * n * sizeof(int16_t), cudaMemcpyDefault, s));
	}
	if (timed) cudaEventRecord (g->ev[6], s);
	g->stats.kernel_launches += 9;
	if (capture.on)
	{
		cudaGraph_t graph = NULL;
		capture.on = false;
		CK (cudaStreamEndCapture (s, &graph));
		cudaError_t ge = cudaGraphInstantiate (&g->graphs[gslot].exec, graph, 0);
		cudaGraphDestroy (graph);
		if (ge != cudaSuccess) { g->graphs[gslot].exec = NULL; snprintf (err, errlen, "cudaGraphInstantiate: %s", cudaGetErrorString (ge)); return RC_ERR_CUDA; }
		memcpy (&g->graphs[gslot].w, &w, sizeof(w)); memcpy (&g->graphs[gslot].f, &f0, sizeof(f0));
		g->graphs[gslot].used = ++g->graph_clock;
		CK (cudaGraphLaunch (g->graphs[gslot].exec, s));
	}
	CK (cudaGetLastError ());
	return RC_OK;
}

static void collect_stats (rc_gpu_t *g, int n, bool timed)
{
	unsigned long long a[16];
	cudaMemcpy (a, g->d_alloc, sizeof(a), cudaMemcpyDeviceToHost);
	g->stats.alias_fragments = (long long)a[9]; g->stats.alias_fragments_front = (long long)a[10]; g->stats.alias_spans = (long long)a[11];
	g->stats.fullcells += (long long)a[4];
	g->stats.slowcells += (long long)a[3];
	g->stats.spans += (long long)(a[0] & 0xFFFFFFFFu);
	g->stats.blocks += (long long)n * g->height * ((g->width + 7) / 8) + (long long)(a[0] & 0xFFFFFFFFu);
	g->stats.cache_jobs += (long long)a[2];
	std::vector<int> ne (n), nf (n);
	cudaMemcpy (ne.data (), g->d_nedges, sizeof(int) * n, cudaMemcpyDeviceToHost);
	cudaMemcpy (nf.data (), g->d_nfaces, sizeof(int) * n, cudaMemcpyDeviceToHost);
	for (int i = 0; i < n; i++) { g->stats.edges += ne[i]; g->stats.faces += nf[i]; }
	if (g->stats.cache_jobs > 0)
	{
		int nj = (int)(a[2] < (unsigned long long)g->jobcap ? a[2] : g->jobcap);
		std::vector<rc_cachejob_t> jobs (nj);
		if (nj) cudaMemcpy (jobs.data (), g->d_jobs, sizeof(rc_cachejob_t) * nj, cudaMemcpyDeviceToHost);
		for (int i = 0; i < nj; i++) g->stats.cache_texels += (long long)jobs[i].width * jobs[i].height;
	}
	if (timed)
	{
		float ms;
		cudaEventElapsedTime (&ms, g->ev[0], g->ev[1]); g->stats.ms_walk += ms;
		cudaEventElapsedTime (&ms, g->ev[1], g->ev[2]); g->stats.ms_faces += ms;
		cudaEventElapsedTime (&ms, g->ev[2], g->ev[3]); g->stats.ms_scan += ms;
		cudaEventElapsedTime (&ms, g->ev[3], g->ev[4]); g->stats.ms_surfprep += ms;
		cudaEventElapsedTime (&ms, g->ev[4], g->ev[5]); g->stats.ms_cache += ms;
		cudaEventElapsedTime (&ms, g->ev[5], g->ev[6]); g->stats.ms_draw += ms;
		cudaEventElapsedTime (&ms, g->ev[5], g->ev[7]); g->stats.ms_spanseg += ms;
		if (cudaEventElapsedTime (&ms, g->ev[7], g->ev[8]) == cudaSuccess) g->stats.ms_cells += ms;
		else cudaGetLastError ();
		cudaEventElapsedTime (&ms, g->ev[0], g->ev[6]); g->stats.ms_total += ms;
		if (g->cur_batch && g->cur_batch->nzres)
		{
			cudaEventElapsedTime (&ms, g->ev[10], g->ev[11]); g->stats.ms_alias += ms;
			cudaEventElapsedTime (&ms, g->ev[11], g->ev[12]); g->stats.ms_particles += ms;
			cudaEventElapsedTime (&ms, g->ev[12], g->ev[13]); g->stats.ms_zresolve += ms;
		}
	}
}

/* deferred mirror: a copy of an earlier batch may still be reading this frame buffer */
static void mirror_wait_pending (rc_gpu_t *g, const void *fb_dev, cudaStream_t s)
{
	for (int i = 0; i < 4; i++)
		if (g->mpend[i].valid && g->mpend[i].fb == fb_dev)
			cudaStreamWaitEvent (s, g->mpend[i].ev, 0);
}

/* deferred mirror: the whole batch goes to the mirror on the copy stream; nothing on `s` waits for it */
static int mirror_enqueue (rc_gpu_t *g, void *fb_dev, void *z_dev, int n, cudaStream_t s, char *err, size_t errlen)
{
	const size_t px = (size_t)g->width * g->height;
	int slot = -1;
	for (int i = 0; i < 4 && slot < 0; i++) if (g->mpend[i].valid && g->mpend[i].fb == fb_dev) slot = i;
	for (int i = 0; i < 4 && slot < 0; i++) if (!g->mpend[i].valid) slot = i;
	if (slot < 0)
	{
		/* more than four frame buffers in rotation: the oldest entry is waited for and reused */
		slot = 0;
		CK (cudaEventSynchronize (g->mpend[0].ev));
	}
	CK (cudaEventRecord (g->evrender, s));
	CK (cudaStreamWaitEvent (g->copystream, g->evrender, 0));
	CK (cudaMemcpyAsync (g->mirror_fb, fb_dev, px * n, cudaMemcpyDefault, g->copystream));
	if (g->mirror_z && z_dev)
		CK (cudaMemcpyAsync (g->mirror_z, z_dev, px * n * sizeof(int16_t), cudaMemcpyDefault, g->copystream));
	CK (cudaEventRecord (g->mpend[slot].ev, g->copystream));
	g->mpend[slot].fb = fb_dev; g->mpend[slot].valid = 1;
	return RC_OK;
}

int rc_gpu_render (rc_gpu_t *g, const rc_batch_t *batch, void *fb_dev, void *z_dev, void *stream,
		   int *status, char *err, size_t errlen)
{
	const rc_view_t *views = batch->views;
	const int n = batch->n;
	/* NULL is the legacy default stream (include/ref_cuda.h): a caller on the default stream is ordered with its own
	 * producers and consumers.  The host-output paths pass the renderer's own stream explicitly. */
	cudaStream_t s = stream ? (cudaStream_t)stream : cudaStreamLegacy;
	const bool deferred = !(g->rflags & RC_RF_NO_MIRROR) && !g->host_fb && g->mirror_fb && g->mirror_deferred;
	const bool mirrored = !(g->rflags & RC_RF_NO_MIRROR) && !g->host_fb && g->mirror_fb && !deferred;
	if (mirrored) { g->host_fb = g->mirror_fb; g->host_z = z_dev ? g->mirror_z : NULL; }
	if (deferred)
		mirror_wait_pending (g, fb_dev, s);
	struct unmirror_t { rc_gpu_t *g; bool on; ~unmirror_t () { if (on) { g->host_fb = NULL; g->host_z = NULL; } } } unmirror = { g, mirrored };
	size_t px = (size_t)g->width * g->height;
	int st = 0;
	if (!g->haveworld) { snprintf (err, errlen, "no world loaded"); return RC_ERR_NOWORLD; }
	cudaSetDevice (g->device);
	memset (&g->stats, 0, sizeof(g->stats));
	if (g->flush_pending)
	{
		do_flush (g, s);
		g->stats.kernel_launches++;
	}
	if (batch->ndlights > g->dlightcap)
	{
		cudaFree (g->d_dlights);
		g->dlightcap = batch->ndlights + 1024;
		CK (cudaMalloc (&g->d_dlights, sizeof(rc_dlight_t) * g->dlightcap));
	}
	if (batch->nwarp > g->warpcap)
	{
		cudaFree (g->d_warpbuf);
		g->warpcap = batch->nwarp;
		CK (cudaMalloc (&g->d_warpbuf, (size_t)g->warpcap * RC_WARP_W * RC_WARP_H));
	}
	if (batch->nbents > g->bentcap)
	{
		cudaFree (g->d_bents); g->bentcap = batch->nbents + 256;
		CK (cudaMalloc (&g->d_bents, sizeof(rc_bent_t) * g->bentcap));
	}
	if (batch->nbents)
		CK (cudaMemcpyAsync (g->d_bents, batch->bents, sizeof(rc_bent_t) * batch->nbents, cudaMemcpyHostToDevice, s));
	if (batch->nalias > g->aliascap)
	{
		cudaFree (g->d_alias); g->aliascap = batch->nalias + 256;
		CK (cudaMalloc (&g->d_alias, sizeof(rc_aliasinst_t) * g->aliascap));
	}
	if (batch->nfinalverts > g->finalvertcap)
	{
		cudaFree (g->d_finalverts); g->finalvertcap = batch->nfinalverts + 4096;
		CK (cudaMalloc (&g->d_finalverts, sizeof(rc_finalvert_t) * (size_t)g->finalvertcap));
	}
	if (batch->nsprites > g->spritecap)
	{
		cudaFree (g->d_sprites); g->spritecap = batch->nsprites + 256;
		CK (cudaMalloc (&g->d_sprites, sizeof(rc_spriteinst_t) * (size_t)g->spritecap));
	}
	if (batch->nsprites)
		CK (cudaMemcpyAsync (g->d_sprites, batch->sprites, sizeof(rc_spriteinst_t) * batch->nsprites, cudaMemcpyHostToDevice, s));
	if (batch->nparticles > g->particlecap)
	{
		cudaFree (g->d_particles); g->particlecap = batch->nparticles + 4096;
		CK (cudaMalloc (&g->d_particles, sizeof(rc_particle_t) * (size_t)g->particlecap));
	}
	if (batch->ncolormaps > g->colormapcap)
	{
		cudaFree (g->d_colormaps); g->colormapcap = batch->ncolormaps;
		CK (cudaMalloc (&g->d_colormaps, (size_t)g->colormapcap * 16384));
	}
	if ((size_t)batch->nzres * px > g->zrescap)
	{
		cudaFree (g->d_zres); g->zrescap = (size_t)batch->nzres * px;
		CK (cudaMalloc (&g->d_zres, sizeof(unsigned long long) * g->zrescap));
	}
	if (batch->nalias)
	{
		/* span queue of the alias rasteriser: sized from what the previous entity batch asked for (a first guess before
		 * that); triangles that find it full are drawn by their own thread, so the size is never a matter of correctness */
		long long want = g->aspan_need + g->aspan_need / 4;
		const long long guess = (long long)(batch->nalias < 32768 ? batch->nalias : 32768) * g->maxtris * 4;
		if (want < guess) want = guess;
		if (want < (1 << 20)) want = 1 << 20;
		if (want > (256ll << 20)) want = 256ll << 20;
		if (want > g->aspancap)
		{
			CK (cudaStreamSynchronize (s));
			cudaFree (g->d_aspans); g->d_aspans = NULL; g->aspancap = 0;
			if (cudaMalloc (&g->d_aspans, sizeof(rc_aspan_t) * (size_t)want) == cudaSuccess) g->aspancap = want;
			else cudaGetLastError ();		/* no queue: direct drawing */
		}
		long long wt = (long long)g->aspantri_need + g->aspantri_need / 4;
		const long long guesst = (long long)(batch->nalias < 32768 ? batch->nalias : 32768) * g->maxtris;
		if (wt < guesst) wt = guesst;
		if (wt > (1ll << 28) - 1) wt = (1ll << 28) - 1;	/* the counter has 28 bits for triangles */
		if (wt > g->aspantricap)
		{
			CK (cudaStreamSynchronize (s));
			cudaFree (g->d_aspantris); g->d_aspantris = NULL; g->aspantricap = 0;
			if (cudaMalloc (&g->d_aspantris, sizeof(rc_aspantri_t) * (size_t)wt) == cudaSuccess) g->aspantricap = (int)wt;
			else cudaGetLastError ();
		}
	}
	if (batch->nalias)
		CK (cudaMemcpyAsync (g->d_alias, batch->alias, sizeof(rc_aliasinst_t) * batch->nalias, cudaMemcpyHostToDevice, s));
	if (batch->nparticles)
		CK (cudaMemcpyAsync (g->d_particles, batch->particles, sizeof(rc_particle_t) * batch->nparticles, cudaMemcpyHostToDevice, s));
	if (batch->nzres)
	{
		CK (cudaMemcpyAsync (g->d_colormaps, batch->colormaps, (size_t)batch->ncolormaps * 16384, cudaMemcpyHostToDevice, s));
		CK (cudaMemsetAsync (g->d_zres, 0, sizeof(unsigned long long) * (size_t)batch->nzres * px, s));
	}
	g->cur_batch = batch;
	if (!z_dev && batch->nzres)
	{
		if (g->zscratchcap < px * n)
		{
			cudaFree (g->d_zscratch); g->zscratchcap = px * n;
			CK (cudaMalloc (&g->d_zscratch, g->zscratchcap * sizeof(int16_t)));
		}
		z_dev = g->d_zscratch;
	}
	if (batch->ndlights)
		CK (cudaMemcpyAsync (g->d_dlights, batch->dlights, sizeof(rc_dlight_t) * batch->ndlights, cudaMemcpyHostToDevice, s));
	for (int attempt = 0; attempt < 4; attempt++)
	{
		int r = ensure_chunk (g, n, batch->nbents ? 512 : 0, err, errlen);
		if (r != RC_OK) return r;
		st = 0;		/* the device status words are cleared by the first chunk's k_reset_chunk */
		for (int first = 0; first < n; first += g->chunkcap)
		{
			int cn = n - first < g->chunkcap ? n - first : g->chunkcap;
			r = run_chunk (g, views + first, first, cn, (uint8_t *)fb_dev + px * first,
				       z_dev ? (int16_t *)z_dev + px * first : NULL, s, g->profiling != 0, err, errlen);
			if (r != RC_OK) return r;
			if (g->profiling || n > g->chunkcap)
			{
				/* d_views is reused by the next chunk and the stats are read here */
				CK (cudaStreamSynchronize (s));
				if (g->profiling) collect_stats (g, cn, true);
			}
		}
		if (g->rflags & RC_RF_DEFER)
			break;		/* one part of a large batch: nothing waits here, rc_gpu_render_finish reads the status */
		if (g->amode)
		{
			/* pipelined call: the status lands in pinned memory, nothing waits here; a pool overflow is
			 * reported by rc_gpu_wait_host and the caller renders that batch again with the blocking call */
			CK (launch_k (k_status_out, dim3 (1), dim3 (1), 0, s, false, (const int *)g->d_status, &g->astatus[g->amode - 1]));
			CK (cudaEventRecord (g->adone_s[g->amode - 1], s));
			break;
		}
		int stw[2] = { 0, 0 };
		unsigned long long qneed[2] = { 0ull, 0ull };
		CK (cudaMemcpyAsync (stw, g->d_status, sizeof(stw), cudaMemcpyDeviceToHost, s));
		if (batch->nalias)
			CK (cudaMemcpyAsync (qneed, g->d_alloc + 6, sizeof(qneed), cudaMemcpyDeviceToHost, s));
		CK (cudaStreamSynchronize (s));
		if (batch->nalias) { g->aspan_need = (long long)qneed[0]; g->aspantri_need = (int)(qneed[1] < (1ull << 28) ? qneed[1] : (1ull << 28) - 1); }
		st = stw[0];
		g->stats.alias_range_events = stw[1];
#ifdef RC_TRACE
		{
			unsigned long long tr[32];
			static const char *nm[10] = { "reset", "front", "scan", "prep", "cache", "spanseg", "draw", "ends", "slow", "flush" };
			cudaMemcpyFromSymbol (tr, g_trace, sizeof(tr));
			const unsigned long long t0 = tr[0];
			fprintf (stderr, "trace(us)");
			for (int k = 0; k < 9; k++)
				if (tr[2 * k + 1]) fprintf (stderr, " %s %.1f-%.1f", nm[k], (double)(long long)(tr[2 * k] - t0) * 1e-3, (double)(long long)(tr[2 * k + 1] - t0) * 1e-3);
			fprintf (stderr, "\n");
		}
#endif
#ifdef RC_FRONT_TIMING
		{
			int tq[16];
			cudaMemcpy (tq, g->d_status, sizeof(tq), cudaMemcpyDeviceToHost);
			fprintf (stderr, "k_front max cycles: reset %d nodetest %d walk %d mark %d list %d | facescan %d edge %d finish %d | fixup %d\n", tq[4], tq[5], tq[6], tq[7], tq[8], tq[9], tq[10], tq[11], tq[12]);
		}
#endif

		if ((st & RC_STAT_AET_OVERFLOW) && !g->scan_large && !g->amode)
		{
			/* more active edges on a scanline (or candidates in an 8-line band) than the small instance of the scan
			 * kernel holds: the large one from now on, and this batch again */
			g->scan_large = 1;
			for (int i = 0; i < RC_GRAPHS; i++) if (g->graphs[i].exec) { cudaGraphExecDestroy (g->graphs[i].exec); g->graphs[i].exec = NULL; g->graphs[i].used = 0; }
			memset (&g->stats, 0, sizeof(g->stats));
			continue;
		}
		if (!(st & (RC_STAT_SPAN_POOL | RC_STAT_CACHE_POOL)))
			break;
		if (st & RC_STAT_SPAN_POOL)
		{
			/* the span / block pool heuristic was too small for this scene: grow and redo */
			g->spans_per_line *= 2;
			free_chunk (g);
		}
		/* surface-cache pool (or table) exhausted, or entries left half-built: D_FlushCaches and redo */
		memset (&g->stats, 0, sizeof(g->stats));
		do_flush (g, s);
	}
	if (deferred)
	{
		const int mr = mirror_enqueue (g, fb_dev, z_dev, n, s, err, errlen);
		if (mr != RC_OK) return mr;
	}
	g->launches_total += g->stats.kernel_launches;
	g->last_n = n < g->chunkcap ? n : g->chunkcap;
	if (status) *status = st;
	return RC_OK;
}

/* Large batches in parts (RC_RenderViewsDevice): the host sets up part k+1 -- per-view constants on all its threads, the
 * upload of the view records -- while the device draws part k.  A part is enqueued like a call of its own but nothing
 * waits for it, the status words accumulate over the parts, the CUDA-graph cache is left alone (every part has its own
 * output pointer) and the mirror copy is left to the end of the whole batch. */
int rc_gpu_can_split (const rc_gpu_t *g)
{
	return !g->profiling && !g->host_fb && !g->groupfn && !g->amode;
}

int rc_gpu_render_begin (rc_gpu_t *g, void *fb_dev, void *stream)
{
	if (g->mirror_fb && g->mirror_deferred)
		mirror_wait_pending (g, fb_dev, stream ? (cudaStream_t)stream : cudaStreamLegacy);
	return RC_OK;
}

int rc_gpu_render_part (rc_gpu_t *g, const rc_batch_t *part, void *fb_part, void *z_part, void *stream, int first_part, char *err, size_t errlen)
{
	g->rflags = RC_RF_DEFER | RC_RF_NO_MIRROR | (first_part ? 0 : RC_RF_KEEP_STATUS);
	const int r = rc_gpu_render (g, part, fb_part, z_part, stream, NULL, err, errlen);
	g->rflags = 0;
	return r;
}

/* the end of a batch drawn in parts: the mirror copy of the whole batch, then the status of all parts (this waits) */
int rc_gpu_render_finish (rc_gpu_t *g, void *fb_dev, void *z_dev, int n, void *stream, int *status, char *err, size_t errlen)
{
	cudaStream_t s = stream ? (cudaStream_t)stream : cudaStreamLegacy;
	const size_t px = (size_t)g->width * g->height;
	if (g->mirror_fb && g->mirror_deferred)
	{
		const int mr = mirror_enqueue (g, fb_dev, z_dev, n, s, err, errlen);
		if (mr != RC_OK) return mr;
	}
	else if (g->mirror_fb)
	{
		CK (cudaMemcpyAsync (g->mirror_fb, fb_dev, px * n, cudaMemcpyDefault, s));
		if (g->mirror_z && z_dev)
			CK (cudaMemcpyAsync (g->mirror_z, z_dev, px * n * sizeof(int16_t), cudaMemcpyDefault, s));
	}
	int stw[2] = { 0, 0 };
	unsigned long long qneed[2] = { 0ull, 0ull };
	CK (cudaMemcpyAsync (stw, g->d_status, sizeof(stw), cudaMemcpyDeviceToHost, s));
	CK (cudaMemcpyAsync (qneed, g->d_alloc + 6, sizeof(qneed), cudaMemcpyDeviceToHost, s));	/* what the parts asked of the alias span queue */
	CK (cudaStreamSynchronize (s));
	if (qneed[0] || qneed[1]) { g->aspan_need = (long long)qneed[0]; g->aspantri_need = (int)(qneed[1] < (1ull << 28) ? qneed[1] : (1ull << 28) - 1); }
	g->stats.alias_range_events = stw[1];
	if (status) *status = stw[0];
	return RC_OK;
}

int rc_gpu_render_host (rc_gpu_t *g, const rc_batch_t *batch, uint8_t *fb_out, int16_t *z_out,
			int *status, char *err, size_t errlen)
{
	const int n = batch->n;
	size_t px = (size_t)g->width * g->height;
	cudaSetDevice (g->device);
	if (g->outcap < px * n)
	{
		cudaFree (g->d_fb); cudaFree (g->d_zb);
		g->d_fb = NULL; g->d_zb = NULL; g->outcap = 0;
		CK (cudaMalloc (&g->d_fb, px * n));
		CK (cudaMalloc (&g->d_zb, px * n * sizeof(int16_t)));
		g->outcap = px * n;
	}
	/* every pixel of a full-screen view is written by the world spans; views that use a
	 * sub-rectangle leave the rest as the caller's buffer had it in the reference, here 0 */
	CK (cudaMemsetAsync (g->d_fb, 0, px * n, g->stream));
	/* the device->host copies are issued by run_chunk (overlapped with the draw kernel where possible);
	 * rc_gpu_render returns when the stream, copies included, has drained */
	g->host_fb = fb_out; g->host_z = z_out;
	int r = rc_gpu_render (g, batch, g->d_fb, g->d_zb, g->stream, status, err, errlen);
	g->host_fb = NULL; g->host_z = NULL;
	return r;
}

/* Pipelined host output: enqueues the batch and returns; at most RC_ASYNC_DEPTH calls in flight.  The device->host
 * copies of this call run on the copy stream while the next call's kernels run, each call drawing into
 * its own device planes. */
int rc_gpu_render_host_async (rc_gpu_t *g, const rc_batch_t *batch, uint8_t *fb_out, int16_t *z_out, char *err, size_t errlen)
{
	const int n = batch->n;
	size_t px = (size_t)g->width * g->height;
	cudaSetDevice (g->device);
	if (g->ainflight >= RC_ASYNC_DEPTH) { snprintf (err, errlen, "%d asynchronous calls already in flight: RC_WaitViews first", RC_ASYNC_DEPTH); return RC_ERR_ARG; }
	if (g->profiling) { snprintf (err, errlen, "asynchronous rendering needs profiling off"); return RC_ERR_ARG; }
	const int slot = (g->ahead + g->ainflight) % RC_ASYNC_DEPTH;
	if (g->aoutcap[slot] < px * n)
	{
		cudaFree (g->d_afb[slot]); cudaFree (g->d_azb[slot]);
		g->d_afb[slot] = NULL; g->d_azb[slot] = NULL; g->aoutcap[slot] = 0;
		CK (cudaMalloc (&g->d_afb[slot], px * n + 8));
		CK (cudaMalloc (&g->d_azb[slot], px * n * sizeof(int16_t)));
		g->aoutcap[slot] = px * n;
	}
	/* cleared by a kernel, not cudaMemsetAsync: nothing of this call may queue on a copy engine ahead of its own
	 * frame-buffer copies (the engine would be busy with the previous call's) */
	CK (launch_k (k_fill64, dim3 (g->numsms * 4), dim3 (256), 0, g->stream, g->use_pdl != 0, (unsigned long long *)g->d_afb[slot], 0ull, (size_t)((px * n + 7) / 8)));
	g->astatus[slot] = 0;
	CK (cudaEventRecord (g->adone_copy[slot], g->copystream));	/* paths that copy on the main stream leave this one as "done" */
	g->host_fb = fb_out; g->host_z = z_out; g->amode = slot + 1;
	int st = 0;
	int r = rc_gpu_render (g, batch, g->d_afb[slot], g->d_azb[slot], g->stream, &st, err, errlen);
	g->host_fb = NULL; g->host_z = NULL; g->amode = 0;
	if (r == RC_OK) g->ainflight++;
	return r;
}

/* waits for the oldest asynchronous call; its buffers are complete when this returns */
int rc_gpu_wait_host (rc_gpu_t *g, int *status, char *err, size_t errlen)
{
	cudaSetDevice (g->device);
	if (g->ainflight <= 0) { snprintf (err, errlen, "no asynchronous call in flight"); return RC_ERR_ARG; }
	const int slot = g->ahead;
	CK (cudaEventSynchronize (g->adone_s[slot]));
	CK (cudaEventSynchronize (g->adone_copy[slot]));
	if (status) *status = g->astatus[slot];
	/* a pool overflow leaves keys without blocks behind (and the allocation pointer past the end): the next render
	 * starts with D_FlushCaches, so the blocking re-render the caller owes this batch -- and every later batch -- is clean;
	 * batches already queued behind this one report RC_STAT_CACHE_POOL themselves (rc_span_seg) */
	if (g->astatus[slot] & (RC_STAT_CACHE_POOL | RC_STAT_SPAN_POOL))
		g->flush_pending = 1;
	g->ahead = (g->ahead + 1) % RC_ASYNC_DEPTH; g->ainflight--;
	return RC_OK;
}

int rc_gpu_present (rc_gpu_t *g, const uint32_t *pal32, int n, const void *fb, void *rgbx, void *stream, int host, char *err, size_t errlen)
{
	cudaSetDevice (g->device);
	cudaStream_t s = host ? g->stream : (stream ? (cudaStream_t)stream : cudaStreamLegacy);	/* NULL: the legacy default stream */
	const size_t px = (size_t)g->width * g->height;
	if (n > g->palcap)
	{
		cudaDeviceSynchronize ();
		cudaFree (g->d_pal32); g->d_pal32 = NULL; g->palcap = 0;
		cudaFreeHost (g->h_pal32[0]); cudaFreeHost (g->h_pal32[1]); g->h_pal32[0] = g->h_pal32[1] = NULL;
		CK (cudaMalloc (&g->d_pal32, sizeof(uint32_t) * 256 * (size_t)(n + 64)));
		CK (cudaHostAlloc (&g->h_pal32[0], sizeof(uint32_t) * 256 * (size_t)(n + 64), cudaHostAllocDefault));
		CK (cudaHostAlloc (&g->h_pal32[1], sizeof(uint32_t) * 256 * (size_t)(n + 64), cudaHostAllocDefault));
		g->palcap = n + 64; g->palvalid = 0;
		cudaEventRecord (g->evpal[0], s); cudaEventRecord (g->evpal[1], s);
	}
	const uint8_t *fbd = (const uint8_t *)fb;
	uint32_t *outd = (uint32_t *)rgbx;
	if (host)
	{
		if (g->presentcap < px * n)
		{
			cudaFree (g->d_presentin); cudaFree (g->d_presentout); g->d_presentin = NULL; g->d_presentout = NULL; g->presentcap = 0;
			CK (cudaMalloc (&g->d_presentin, px * n));
			CK (cudaMalloc (&g->d_presentout, px * n * sizeof(uint32_t)));
			g->presentcap = px * n;
		}
		CK (cudaMemcpyAsync (g->d_presentin, fb, px * n, cudaMemcpyHostToDevice, s));
		fbd = g->d_presentin; outd = g->d_presentout;
	}
	/* palettes go through a pinned two-slot ring, so the upload is a true asynchronous copy (a pageable source costs the
	 * stream tens of microseconds); a slot is reused only after the copy that read it has completed */
	if (!pal32)
	{
		if (n > g->palvalid) { snprintf (err, errlen, "no palettes from a previous call for %d views", n); return RC_ERR_ARG; }
	}
	else
	{
		const int slot = g->palslot; g->palslot ^= 1;
		g->palvalid = n;
		CK (cudaEventSynchronize (g->evpal[slot]));
		memcpy (g->h_pal32[slot], pal32, sizeof(uint32_t) * 256 * (size_t)n);
		CK (cudaMemcpyAsync (g->d_pal32, g->h_pal32[slot], sizeof(uint32_t) * 256 * (size_t)n, cudaMemcpyHostToDevice, s));
		CK (cudaEventRecord (g->evpal[slot], s));
	}
	int bx = (int)((px / 4 + 256 * 4 - 1) / (256 * 4));
	const int cap = (g->numsms * 8 + n - 1) / n;		/* about 8 resident CTAs per SM over the whole batch */
	if (bx > cap) bx = cap;
	if (bx < 1) bx = 1;
	k_present<<<dim3 (bx, n), 256, 0, s>>> (fbd, outd, g->d_pal32, (unsigned)px);
	CK (cudaGetLastError ());
	if (host)
	{
		CK (cudaMemcpyAsync (rgbx, g->d_presentout, px * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		CK (cudaStreamSynchronize (s));
	}
	return RC_OK;
}

int rc_gpu_upload_draw2d (rc_gpu_t *g, const uint8_t *blob, size_t len, const void *pics, int npics, char *err, size_t errlen)
{
	cudaSetDevice (g->device);
	cudaDeviceSynchronize ();
	cudaFree (g->d_d2blob); cudaFree (g->d_d2pics); g->d_d2blob = NULL; g->d_d2pics = NULL;
	CK (cudaMalloc (&g->d_d2blob, len + 16));
	CK (cudaMalloc (&g->d_d2pics, sizeof(rc_pic_t) * (size_t)(npics + 1)));
	CK (cudaMemcpy (g->d_d2blob, blob, len, cudaMemcpyHostToDevice));
	CK (cudaMemcpy (g->d_d2pics, pics, sizeof(rc_pic_t) * (size_t)npics, cudaMemcpyHostToDevice));
	return RC_OK;
}

int rc_gpu_draw2d (rc_gpu_t *g, const rc_draw2d_t *cmds, int ncmds, const int *first, int n, void *fb, void *stream, int host, char *err, size_t errlen)
{
	cudaSetDevice (g->device);
	cudaStream_t s = host ? g->stream : (stream ? (cudaStream_t)stream : cudaStreamLegacy);	/* NULL: the legacy default stream */
	const size_t px = (size_t)g->width * g->height;
	if (!g->d_d2blob) { snprintf (err, errlen, "2D assets not uploaded"); return RC_ERR_ARG; }
	if (ncmds > g->d2cmdcap || n + 1 > g->d2firstcap)
	{
		cudaStreamSynchronize (s);
		cudaFree (g->d_d2cmds); cudaFree (g->d_d2first); g->d_d2cmds = NULL; g->d_d2first = NULL;
		g->d2cmdcap = ncmds + 1024; g->d2firstcap = n + 1 + 256;
		CK (cudaMalloc (&g->d_d2cmds, sizeof(rc_draw2d_t) * (size_t)g->d2cmdcap));
		CK (cudaMalloc (&g->d_d2first, sizeof(int) * (size_t)g->d2firstcap));
	}
	if (ncmds) CK (cudaMemcpyAsync (g->d_d2cmds, cmds, sizeof(rc_draw2d_t) * (size_t)ncmds, cudaMemcpyHostToDevice, s));
	CK (cudaMemcpyAsync (g->d_d2first, first, sizeof(int) * (size_t)(n + 1), cudaMemcpyHostToDevice, s));
	uint8_t *fbd = (uint8_t *)fb;
	if (host)
	{
		if (g->presentcap < px * n)
		{
			cudaFree (g->d_presentin); cudaFree (g->d_presentout); g->d_presentin = NULL; g->d_presentout = NULL; g->presentcap = 0;
			CK (cudaMalloc (&g->d_presentin, px * n));
			CK (cudaMalloc (&g->d_presentout, px * n * sizeof(uint32_t)));
			g->presentcap = px * n;
		}
		CK (cudaMemcpyAsync (g->d_presentin, fb, px * n, cudaMemcpyHostToDevice, s));
		fbd = g->d_presentin;
	}
	rc_d2ctx_t x;
	x.blob = g->d_d2blob; x.pics = (const rc_pic_t *)g->d_d2pics; x.width = g->width; x.height = g->height;
	k_draw2d<<<n, 256, 0, s>>> (x, g->d_d2cmds, g->d_d2first, fbd);
	CK (cudaGetLastError ());
	if (host)
	{
		CK (cudaMemcpyAsync (fb, g->d_presentin, px * n, cudaMemcpyDeviceToHost, s));
		CK (cudaStreamSynchronize (s));
	}
	else
		CK (cudaStreamSynchronize (s));		/* cmds / first are the caller's: they may be reused when this returns */
	return RC_OK;
}

void rc_gpu_get_stats (rc_gpu_t *g, rc_stats_t *out) { *out = g->stats; }
void rc_gpu_set_profiling (rc_gpu_t *g, int on) { g->profiling = on; }

int rc_gpu_debug_read (rc_gpu_t *g, int kind, int view, void *dst, int cap)
{
	cudaSetDevice (g->device);
	cudaDeviceSynchronize ();
	if (!g->d_views || view < 0 || view >= g->last_n) return RC_ERR_ARG;
	if (kind == 0)
	{
		int n = 0;
		cudaMemcpy (&n, g->d_nedges + view, sizeof(int), cudaMemcpyDeviceToHost);
		if (n > g->edgecap) n = g->edgecap;
		if (n > cap) n = cap;
		cudaMemcpy (dst, g->d_edges + (size_t)view * g->edgecap, sizeof(rc_edge_t) * n, cudaMemcpyDeviceToHost);
		return n;
	}
	if (kind == 1)
	{
		int n = 0;
		cudaMemcpy (&n, g->d_nfaces + view, sizeof(int), cudaMemcpyDeviceToHost);
		n += 2;
		if (n > cap) n = cap;
		cudaMemcpy (dst, g->d_surfs + (size_t)view * g->surfcap, sizeof(rc_surf_t) * n, cudaMemcpyDeviceToHost);
		return n;
	}
	if (kind == 2)
	{
		unsigned long long a = 0;
		cudaMemcpy (&a, g->d_alloc, sizeof(a), cudaMemcpyDeviceToHost);
		int n = (int)(a & 0xFFFFFFFFu);
		if (n > g->spancap) n = g->spancap;
		if (n > cap) n = cap;
		cudaMemcpy (dst, g->d_spans, sizeof(rc_span_t) * n, cudaMemcpyDeviceToHost);
		return n;
	}
	return RC_ERR_ARG;
}

} /* extern "C" */
