/*
 * sonde_b200.cu — the extern "C" shim declared in include/sonde_b200.h.
 *
 * Owns the per-channel state in HBM, sorts channels into type-homogeneous CTA groups,
 * and enqueues K1 (demod.cu) + K2 (frame.cu) per process call on one CUDA stream.
 * There is no CPU fallback: without an sm_100 device create() fails with
 * SONDE_ERR_NODEVICE.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "device_state.h"

extern "C" cudaError_t sonde_upload_modems(const sonde_modem *m);
extern "C" cudaError_t sonde_upload_modems_frame(const sonde_modem *m);
extern "C" cudaError_t sonde_upload_gf_tables(void);
extern "C" cudaError_t sonde_launch_demod_gfsk(const demod_params *p, int group_base, int n_groups, int phases,
                                               cudaStream_t stream);
extern "C" cudaError_t sonde_upload_modems_pipe(const sonde_modem *m);
extern "C" cudaError_t sonde_upload_modems_afsk_pipe(const sonde_modem *m);
extern "C" cudaError_t sonde_launch_demod_pipe_afsk(const demod_params *p, int group_base, int n_groups, int layout,
                                                    cudaStream_t stream);
extern "C" cudaError_t sonde_launch_demod_pipe(const demod_params *p, int group_base, int n_groups, int phases,
                                              cudaStream_t stream);
extern "C" cudaError_t sonde_launch_demod_afsk(const demod_params *p, int group_base, int n_groups,
                                               cudaStream_t stream);
extern "C" cudaError_t sonde_launch_frames(const frame_params *p, cudaStream_t stream);
extern "C" cudaError_t sonde_launch_auto_classify(const void *d_in, size_t row_stride, int len, int is_iq,
                                                  const int32_t *d_rows, int n, uint32_t *d_mask, cudaStream_t stream);

struct sonde_b200 {
	sonde_b200_config cfg;
	std::vector<int32_t> types;          /* per VIRTUAL channel (AUTO channels expand to seven)            */
	int n_user = 0;                      /* channels as the caller sees them                               */
	std::vector<int32_t> user_types;     /* as given, SONDE_AUTO allowed                                   */
	std::vector<int32_t> slot0;          /* first virtual channel of each user channel                     */
	std::vector<int32_t> locked;         /* decoder type a user channel reports (SONDE_AUTO = undetermined)*/
	std::vector<int32_t> gchan_host;     /* host copy of d_group_chan                                      */
	int32_t *d_in_row = nullptr, *d_active = nullptr;
	std::vector<int32_t> in_row_host;    /* input row of every (virtual) channel */
	int box_rows_v[4] = {0, 0, 0, 0};    /* per kernel variant: rows of the TMA box (= its group size) when the channels of every
	                                        group sit a fixed number of input rows apart (row_step_v), else 0: one bulk copy per row */
	int row_step_v[4] = {1, 1, 1, 1};
	std::vector<sonde_frame_rec> h_recs; /* staging for fetch() when virtual != user channels              */
	std::vector<int32_t> h_vcounts;
	bool has_auto = false;
	sonde_modem modems[SONDE_NTYPES_];
	int device = 0, n_sms = 0;
	cudaStream_t pstream[3] = {nullptr, nullptr, nullptr};   /* extra copy streams of sonde_b200_process_iq_peer */
	cudaEvent_t ev_piece[3] = {nullptr, nullptr, nullptr};
	unsigned long long peers_enabled = 0;  /* devices this handle's device has been given peer access to */
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};

	/* CTA groups, sorted: GFSK 1-phase | GFSK 2-phase | AFSK */
	int n_groups = 0, groups_v[4] = {0, 0, 0, 0};   /* per kernel variant, see sonde_launch_demod_pipe */
	int group_capacity = 0;                         /* groups allocated at create (all virtual channels active) */
	int32_t *d_group_chan = nullptr, *d_group_type = nullptr, *d_types = nullptr;

	demod_state *d_demod = nullptr;
	afsk_state *d_afsk = nullptr;
	framer_state *d_framer = nullptr;
	uint8_t *d_ring = nullptr;
	uint32_t ring_bytes = 0;
	/* results and host-input staging are double buffered by call parity so that the H2D copy of call
	 * i+1 overlaps the kernels of call i and fetch() of call i can run while call i+1 computes */
	sonde_frame_rec *d_recs[2] = {nullptr, nullptr};
	int32_t *d_counts[2] = {nullptr, nullptr};
	cudaStream_t cstream = nullptr, dstream = nullptr;        /* H2D copies, D2H fetches */
	cudaStream_t fstream = nullptr;                          /* framer kernels when SONDE_FRAME_OVERLAP=1: frame(i) beside demod(i+1) */
	cudaStream_t vstream[4] = {nullptr, nullptr, nullptr, nullptr};   /* one per demod kernel variant: they run concurrently */
	cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t ev_demod[2] = {nullptr, nullptr}, evf[2] = {nullptr, nullptr};
	uint64_t *d_nbits[2] = {nullptr, nullptr};
	cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
	long n_issued = 0, n_fetched = 0, n_truncated = 0;
	float *d_soft = nullptr;
	int max_frames = 0, soft_stride = 0, bits_stride = 0;
	bool rate_registered = false;
	/* AUTO pre-classifier (opt-in, cfg.reserved bit 3): per user channel the set of decoder types
	 * still tried while it is unlocked, when it was narrowed, and the device scratch of the classifier */
	std::vector<uint32_t> plausible;     /* bit t: type t is tried; all ones = the reference's try-all              */
	std::vector<int64_t> narrowed_at;    /* samples processed when `plausible` was narrowed, -1 = not narrowed     */
	bool classify_pending = false;       /* some unlocked AUTO channel has not been classified yet                  */
	int64_t samples_seen = 0;
	int32_t *d_cls_rows = nullptr;
	uint32_t *d_cls_mask = nullptr;
	void *d_in[2] = {nullptr, nullptr};   /* staging for the host-buffer entry points */
	void *d_in16[2] = {nullptr, nullptr}; /* raw int16 IQ staging of sonde_b200_process_iq_s16 */
	int32_t *h_counts = nullptr;     /* pinned */
	long long *d_prof = nullptr;     /* diagnostics: per-CTA stall counters of the pipeline kernel */

	int chunk_index = 0;
	uint64_t bits_before_last = 0;
	long launches = 0;
	bool have_timing = false;
	std::string err;
};

#define SONDE_PEER_STREAMS 4

namespace {

int fail(sonde_b200 *h, int code, const char *what, cudaError_t e = cudaSuccess)
{
	if (h) {
		h->err = what;
		if (e != cudaSuccess) {
			h->err += ": ";
			h->err += cudaGetErrorString(e);
		}
	}
	return code;
}

#define CK(call)                                                        \
	do {                                                                \
		cudaError_t e_ = (call);                                        \
		if (e_ != cudaSuccess) return fail(h, SONDE_ERR_CUDA, #call, e_); \
	} while (0)

/* live handles per device and the sample rate their (per-device, __constant__) modem tables were built for;
 * delta = +1 registers (false if the device is in use at another rate), -1 unregisters */
bool rate_registry(int device, int samplerate, int delta)
{
	static std::mutex mtx;
	static std::map<int, std::pair<int, int>> live;      /* device -> (samplerate, handles) */
	std::lock_guard<std::mutex> lock(mtx);
	auto &e = live[device];
	if (delta > 0) {
		if (e.second > 0 && e.first != samplerate) return false;
		e.first = samplerate;
		e.second++;
	} else if (e.second > 0) {
		e.second--;
	}
	return true;
}

/* Sort the (active) virtual channels into type-homogeneous CTA groups of DEMOD_G, ordered by kernel variant:
 * GFSK 1-phase ~10 slots/symbol | GFSK 1-phase ~20 slots/symbol | GFSK 2-phase | AFSK.  Sets h->groups_v / n_groups.
 * Called at create (everything active) and again whenever AUTO channels lock, so that the surviving virtual
 * channels are packed densely: a CTA costs the same whether one or eight of its lanes carry a channel. */
void build_groups(sonde_b200 *h, const std::vector<int32_t> *active, std::vector<int32_t> &gchan, std::vector<int32_t> &gtype)
{
	const int C = h->cfg.n_channels;
	gchan.clear();
	gtype.clear();
	/* Channels per CTA.  The serial warps of a CTA cost the same for one channel as for eight and the parallel work
	 * scales with the channel count, so the batch is spread over all SMs: the fewest waves of one CTA per SM that
	 * hold it at <= DEMOD_G channels each, then the smallest group size that still fits in those waves
	 * (1024 channels on 148 SMs: 147 CTAs of 7 instead of 128 of 8). */
	/* A batch that mixes kernel variants launches each variant as clusters of two CTAs (launch_tpc_pairs: TPC siblings
	 * then run the same code), so every variant occupies an even number of SMs; the group size is the smallest one whose
	 * padded CTA counts still fit the waves. */
	auto variant_of = [](const sonde_modem &m) { return !m.baud ? -1 : m.afsk ? 3 : m.num_phases == 2 ? 2 : m.freq0 >= 0.19f ? 0 : 1; };
	int gsz_v[4] = {DEMOD_G, DEMOD_G, DEMOD_G, DEMOD_G};
	{
		int per_type[SONDE_NTYPES] = {0};
		for (int c = 0; c < C; c++)
			if (!active || (*active)[c]) per_type[h->types[c]]++;
		int n_var = 0, seen[4] = {0, 0, 0, 0};
		for (int t = 0; t < SONDE_NTYPES; t++)
			if (per_type[t] && variant_of(h->modems[t]) >= 0 && !seen[variant_of(h->modems[t])]++) n_var++;
		const int sms = h->n_sms > 0 ? h->n_sms : 148;
		auto ctas_needed = [&](const int (&g)[4]) {
			int per_var[4] = {0, 0, 0, 0}, total = 0;
			for (int t = 0; t < SONDE_NTYPES; t++) {
				const int v = variant_of(h->modems[t]);
				if (per_type[t] && v >= 0) per_var[v] += (per_type[t] + g[v] - 1) / g[v];
			}
			for (int v = 0; v < 4; v++) total += n_var > 1 ? (per_var[v] + 1) & ~1 : per_var[v];
			return total;
		};
		const int budget = sms * std::max(1, (ctas_needed(gsz_v) + sms - 1) / sms);
		for (int g = 1; g < DEMOD_G; g++) {
			const int u[4] = {g, g, g, g};
			if (ctas_needed(u) <= budget) { for (int v = 0; v < 4; v++) gsz_v[v] = g; break; }
		}
		/* SMs the uniform size leaves over go to the variants whose CTA time grows fastest with its channel count: the
		 * AFSK kernel (its parallel warps are bound by the special-function pipe), then the two-branch M10/M20 one */
		for (bool changed = true; changed;) {
			changed = false;
			for (int v : {3, 2}) {
				if (!seen[v] || gsz_v[v] <= 1) continue;
				int u[4] = {gsz_v[0], gsz_v[1], gsz_v[2], gsz_v[3]};
				u[v]--;
				if (ctas_needed(u) <= budget) { gsz_v[v]--; changed = true; break; }
			}
		}
	}
	auto add_groups = [&](auto pred) {
		int added = 0;
		for (int t = 0; t < SONDE_NTYPES; t++) {
			if (!pred(h->modems[t])) continue;
			const int gsz = gsz_v[variant_of(h->modems[t])];
			std::vector<int32_t> ch;
			for (int c = 0; c < C; c++)
				if (h->types[c] == t && (!active || (*active)[c])) ch.push_back(c);
			for (size_t i = 0; i < ch.size(); i += gsz) {
				for (int k = 0; k < DEMOD_G; k++) gchan.push_back(k < gsz && i + k < ch.size() ? ch[i + k] : -1);
				gtype.push_back(t);
				added++;
			}
		}
		return added;
	};
	/* NCO slots per symbol = 2 / freq0 */
	h->groups_v[0] = add_groups([](const sonde_modem &m) { return m.baud && !m.afsk && m.num_phases == 1 && m.freq0 >= 0.19f; });
	h->groups_v[1] = add_groups([](const sonde_modem &m) { return m.baud && !m.afsk && m.num_phases == 1 && m.freq0 < 0.19f; });
	h->groups_v[2] = add_groups([](const sonde_modem &m) { return m.baud && !m.afsk && m.num_phases == 2; });
	h->groups_v[3] = add_groups([](const sonde_modem &m) { return m.baud && m.afsk; });
	h->n_groups = (int)gtype.size();
	/* tensor staging: a group can be fetched with one tensor copy when its channels are a fixed number of input rows apart —
	 * 1 for a typed batch or AUTO channels, k when k sonde types alternate from channel to channel — and that step is the
	 * same for all groups of the kernel variant */
	int gi = 0;
	for (int v = 0; v < 4; v++) {
		int step = 0;                                        /* 0: not seen yet, -1: no common step */
		for (int k = 0; k < h->groups_v[v]; k++, gi++) {
			const int32_t *ch = &gchan[(size_t)gi * DEMOD_G];
			for (int j = 1; j < DEMOD_G && ch[j] >= 0 && step >= 0 && !h->in_row_host.empty(); j++) {
				const int d = h->in_row_host[ch[j]] - h->in_row_host[ch[j - 1]];
				if (d <= 0 || (step > 0 && d != step)) step = -1;
				else step = d;
			}
		}
		if (step == 0) step = 1;                             /* single-channel groups */
		h->box_rows_v[v] = (h->groups_v[v] > 0 && step > 0 && !h->in_row_host.empty()) ? gsz_v[v] : 0;
		h->row_step_v[v] = step > 0 ? step : 1;
	}
}

uint32_t next_pow2(uint32_t v)
{
	uint32_t p = 1;
	while (p < v) p <<= 1;
	return p;
}

/* most bits one call of `len` samples can add to a channel's stream: the NCO increment is
 * clamped to centre + max_fdev (timing.c:73-75), two NCO counts per bit */
int max_new_bits(const sonde_modem &m, int len)
{
	const double per_slot = ((double)m.freq0 + (double)m.max_fdev) / 2.0;
	return (int)std::ceil((double)len * m.num_phases * per_slot) + 2;
}

}  // namespace

extern "C" {

/* Host-only view of how a batch would be laid out on the GPU (no device needed; what build_groups decides): per kernel
 * variant v = 0..3 (RS41-like, 2400-baud GFSK, two-branch M10/M20, AFSK) out[4 v + {0,1,2,3}] = CTA groups, channels per
 * group, rows of the TMA box (0 = one bulk copy per row), input-row step between the channels of a group; out[16] = CTAs
 * including the padding of the cluster-of-two launches of a mixed batch.  types[c] may be SONDE_AUTO. */
int sonde_b200_debug_plan(const int32_t *types, int n_channels, int samplerate, int n_sms, int32_t *out17)
{
	if (!types || n_channels <= 0 || !out17) return SONDE_ERR_ARG;
	static const int32_t kAutoOrder[SONDE_NTYPES] = {SONDE_RS41, SONDE_M10, SONDE_IMS100, SONDE_DFM09, SONDE_IMET4,
	                                                SONDE_C50, SONDE_MRZN1};
	sonde_b200 tmp;
	sonde_b200 *h = &tmp;
	for (int c = 0; c < n_channels; c++) {
		if (types[c] == SONDE_AUTO) {
			for (int k = 0; k < SONDE_NTYPES; k++) { h->types.push_back(kAutoOrder[k]); h->in_row_host.push_back(c); }
		} else if (types[c] >= 0 && types[c] < SONDE_NTYPES) {
			h->types.push_back(types[c]);
			h->in_row_host.push_back(c);
		} else {
			return SONDE_ERR_ARG;
		}
	}
	h->n_user = n_channels;
	h->cfg.n_channels = (int32_t)h->types.size();
	h->n_sms = n_sms;
	for (int t = 0; t < SONDE_NTYPES; t++)
		if (sonde_modem_init(&h->modems[t], t, samplerate)) memset(&h->modems[t], 0, sizeof(sonde_modem));
	std::vector<int32_t> gchan, gtype;
	build_groups(h, nullptr, gchan, gtype);
	int n_var = 0, total = 0;
	for (int v = 0; v < 4; v++) n_var += h->groups_v[v] > 0;
	for (int v = 0; v < 4; v++) {
		int gsz = 0;
		int gi = 0;
		for (int k = 0; k < v; k++) gi += h->groups_v[k];
		if (h->groups_v[v])
			for (int j = 0; j < DEMOD_G && gchan[(size_t)gi * DEMOD_G + j] >= 0; j++) gsz++;
		out17[4 * v + 0] = h->groups_v[v];
		out17[4 * v + 1] = gsz;
		out17[4 * v + 2] = h->box_rows_v[v];
		out17[4 * v + 3] = h->row_step_v[v];
		total += n_var > 1 ? (h->groups_v[v] + 1) & ~1 : h->groups_v[v];
	}
	out17[16] = total;
	return SONDE_OK;
}

const char *sonde_b200_version(void) { return "sonde_b200 0.1 (sm_100a)"; }

const char *sonde_b200_last_error(const sonde_b200 *h) { return h ? h->err.c_str() : "null handle"; }

int sonde_b200_create(sonde_b200 **out, const sonde_b200_config *cfg)
{
	if (!out) return SONDE_ERR_ARG;
	*out = nullptr;
	if (!cfg || cfg->n_channels <= 0 || cfg->samplerate <= 0 || cfg->max_chunk_len <= 0 || !cfg->types)
		return SONDE_ERR_ARG;
	for (int c = 0; c < cfg->n_channels; c++)
		if (cfg->types[c] != SONDE_AUTO && (cfg->types[c] < 0 || cfg->types[c] >= SONDE_NTYPES)) return SONDE_ERR_ARG;

	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
		return SONDE_ERR_NODEVICE;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10)
		return SONDE_ERR_NODEVICE;       /* kernels are built for sm_100a only */

	sonde_b200 *h = new sonde_b200();
	h->cfg = *cfg;
	/* virtual channels: a fixed channel is itself; an AUTO channel is the seven decoders in the order of
	 * SD/decode.c:174-224, all fed from the same input row */
	static const int32_t kAutoOrder[SONDE_NTYPES] = {SONDE_RS41, SONDE_M10, SONDE_IMS100, SONDE_DFM09, SONDE_IMET4,
	                                                SONDE_C50, SONDE_MRZN1};
	h->n_user = cfg->n_channels;
	h->user_types.assign(cfg->types, cfg->types + cfg->n_channels);
	std::vector<int32_t> in_row;
	for (int c = 0; c < cfg->n_channels; c++) {
		h->slot0.push_back((int32_t)h->types.size());
		if (cfg->types[c] == SONDE_AUTO) {
			h->has_auto = true;
			for (int k = 0; k < SONDE_NTYPES; k++) { h->types.push_back(kAutoOrder[k]); in_row.push_back(c); }
		} else {
			h->types.push_back(cfg->types[c]);
			in_row.push_back(c);
		}
	}
	h->in_row_host = in_row;
	h->locked = h->user_types;
	h->plausible.assign(cfg->n_channels, (1u << SONDE_NTYPES) - 1u);
	h->narrowed_at.assign(cfg->n_channels, -1);
	h->classify_pending = h->has_auto;
	h->cfg.n_channels = (int32_t)h->types.size();        /* device-side channel count from here on */
	h->cfg.types = h->types.data();
	h->device = cfg->device;
	h->n_sms = prop.multiProcessorCount;
	if (h->cfg.fm_gain == 0.0f) h->cfg.fm_gain = 0.636619747f;

	auto bail = [&](int code) { sonde_b200_destroy(h); return code; };

	for (int t = 0; t < SONDE_NTYPES; t++)
		if (sonde_modem_init(&h->modems[t], t, cfg->samplerate)) {
			/* a type that does not exist at this rate is only an error if a channel uses it */
			for (size_t c = 0; c < h->types.size(); c++)
				if (h->types[c] == t) return bail(SONDE_ERR_ARG);
			memset(&h->modems[t], 0, sizeof(sonde_modem));
		}

	if (cudaSetDevice(h->device) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	/* The modem tables live in __constant__ memory, one copy per device: all live handles of a device must agree on
	 * the sample rate they were derived from. */
	if (!rate_registry(h->device, cfg->samplerate, +1)) return bail(SONDE_ERR_STATE);
	h->rate_registered = true;
	if (sonde_upload_modems(h->modems) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	if (sonde_upload_modems_frame(h->modems) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	if (sonde_upload_modems_pipe(h->modems) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	if (sonde_upload_modems_afsk_pipe(h->modems) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	if (sonde_upload_gf_tables() != cudaSuccess) return bail(SONDE_ERR_CUDA);
	/* Stream priorities (matter only with SONDE_FRAME_OVERLAP=1, see run_chunk): frame(i) (fstream) and demod(i+1) (stream)
	 * become runnable at the same moment, when demod(i) retires.  The demodulator is the critical path and needs one large CTA on every SM; if the framer's small CTAs are
	 * placed first, several of them land on each SM and the demodulator's CTA has to wait until they have finished
	 * (measured: the step cost demod + frame although the two overlap).  With the demodulator streams at the highest
	 * priority its CTAs are placed first and the framer fills what is left beside them. */
	int prio_lo = 0, prio_hi = 0;
	if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	{
		const char *e = getenv("SONDE_STREAM_PRIO");              /* experiment switch: 0 = all streams equal */
		if (e && atoi(e) == 0) prio_lo = prio_hi = 0;
	}
	if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	if (cudaStreamCreateWithFlags(&h->cstream, cudaStreamNonBlocking) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	if (cudaStreamCreateWithFlags(&h->dstream, cudaStreamNonBlocking) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	if (cudaStreamCreateWithPriority(&h->fstream, cudaStreamNonBlocking, prio_lo) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	for (int v = 0; v < 4; v++)
		if (cudaStreamCreateWithPriority(&h->vstream[v], cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
		    cudaEventCreateWithFlags(&h->ev_join[v], cudaEventDisableTiming) != cudaSuccess)
			return bail(SONDE_ERR_CUDA);
	for (int b = 0; b < 2; b++)
		if (cudaEventCreateWithFlags(&h->ev_demod[b], cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreate(&h->evf[b]) != cudaSuccess)
			return bail(SONDE_ERR_CUDA);
	for (auto &e : h->ev)
		if (cudaEventCreate(&e) != cudaSuccess) return bail(SONDE_ERR_CUDA);
	for (int b = 0; b < 2; b++)
		if (cudaEventCreateWithFlags(&h->ev_copied[b], cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreateWithFlags(&h->ev_done[b], cudaEventDisableTiming) != cudaSuccess)
			return bail(SONDE_ERR_CUDA);

	/* ---- channel groups: type-homogeneous CTAs, ordered by kernel variant ---------------- */
	const int C = h->cfg.n_channels;
	std::vector<int32_t> gchan, gtype;
	build_groups(h, nullptr, gchan, gtype);
	/* the group size shrinks when fewer channels are active (AUTO locks), so a regrouped batch can have MORE groups
	 * than the all-active one; one group per channel is the bound */
	h->group_capacity = std::max((int)gtype.size(), C);

	/* ---- sizes ----------------------------------------------------------------------------- */
	int bits_max = 0, frames_max = 0;
	uint32_t ring_need = 0;
	for (int t = 0; t < SONDE_NTYPES; t++) {
		if (std::find(h->types.begin(), h->types.end(), t) == h->types.end()) continue;
		const sonde_modem &m = h->modems[t];
		const int nb = max_new_bits(m, cfg->max_chunk_len);
		bits_max = std::max(bits_max, nb);
		/* a window normally consumes a whole frame; iMet-4's framer_adjust (imet4.c:91-117) advances by as little as one
		 * 5-byte subframe of 10-bit characters, so a garbled stream can yield a record every ~50 bits */
		const int min_advance = (t == SONDE_IMET4) ? 50 : m.frame_bits;
		frames_max = std::max(frames_max, nb / min_advance + 2);
		/* two calls are in flight on the ring: frame(i) on fstream still reads while demod(i+1) appends (only
		 * demod(i+2) waits for frame(i), run_chunk), so the span is the framer's 2F + S backlog plus TWO calls' bits */
		ring_need = std::max(ring_need, (uint32_t)((2 * m.frame_bits + m.sync_len + 2 * nb) / 8 + 64));
	}
	h->ring_bytes = next_pow2(ring_need);
	h->max_frames = frames_max;
	h->soft_stride = bits_max;
	h->bits_stride = (bits_max + 7) / 8 + 1;

#define CKB(call) do { if ((call) != cudaSuccess) { h->err = #call; return bail(SONDE_ERR_CUDA); } } while (0)
	CKB(cudaMalloc(&h->d_group_chan, (size_t)h->group_capacity * DEMOD_G * sizeof(int32_t)));
	CKB(cudaMalloc(&h->d_group_type, (size_t)h->group_capacity * sizeof(int32_t)));
	CKB(cudaMalloc(&h->d_types, C * sizeof(int32_t)));
	CKB(cudaMemcpy(h->d_group_chan, gchan.data(), gchan.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
	CKB(cudaMemcpy(h->d_group_type, gtype.data(), gtype.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
	CKB(cudaMemcpy(h->d_types, h->types.data(), C * sizeof(int32_t), cudaMemcpyHostToDevice));
	h->gchan_host = gchan;
	CKB(cudaMalloc(&h->d_in_row, C * sizeof(int32_t)));
	CKB(cudaMemcpy(h->d_in_row, in_row.data(), C * sizeof(int32_t), cudaMemcpyHostToDevice));
	{
		std::vector<int32_t> ones(C, 1);
		CKB(cudaMalloc(&h->d_active, C * sizeof(int32_t)));
		CKB(cudaMemcpy(h->d_active, ones.data(), C * sizeof(int32_t), cudaMemcpyHostToDevice));
	}
	if (h->has_auto) {
		h->h_recs.resize((size_t)C * h->max_frames);
		h->h_vcounts.resize((size_t)C * 2);
	}

	CKB(cudaMalloc(&h->d_demod, (size_t)C * sizeof(demod_state)));
	CKB(cudaMalloc(&h->d_framer, (size_t)C * sizeof(framer_state)));
	CKB(cudaMalloc(&h->d_ring, (size_t)C * h->ring_bytes));
	for (int b = 0; b < 2; b++) {
		CKB(cudaMalloc(&h->d_recs[b], (size_t)C * h->max_frames * sizeof(sonde_frame_rec)));
		CKB(cudaMalloc(&h->d_counts[b], (size_t)C * 2 * sizeof(int32_t)));
		CKB(cudaMalloc(&h->d_nbits[b], (size_t)C * sizeof(uint64_t)));
		CKB(cudaMemset(h->d_nbits[b], 0, (size_t)C * sizeof(uint64_t)));
		CKB(cudaMemset(h->d_counts[b], 0, (size_t)C * 2 * sizeof(int32_t)));
	}
	CKB(cudaMallocHost(&h->h_counts, (size_t)C * 2 * sizeof(int32_t)));
	CKB(cudaMemset(h->d_ring, 0, (size_t)C * h->ring_bytes));
	CKB(cudaMemset(h->d_framer, 0, (size_t)C * sizeof(framer_state)));
	if (cfg->keep_soft) CKB(cudaMalloc(&h->d_soft, (size_t)C * h->soft_stride * sizeof(float)));
	if (h->groups_v[3]) {
		CKB(cudaMalloc(&h->d_afsk, (size_t)C * sizeof(afsk_state)));
		CKB(cudaMemset(h->d_afsk, 0, (size_t)C * sizeof(afsk_state)));
	}

	/* initial demodulator state: agc_init (agc.c:12-16), timing_init (timing.c:14-25), zero filter memory */
	{
		std::vector<demod_state> st(C);
		memset(st.data(), 0, st.size() * sizeof(demod_state));
		for (int c = 0; c < C; c++) {
			st[c].agc_avg = 5.0f;
			st[c].t_freq = h->modems[h->types[c]].freq0;
			st[c].t_state = 1;
		}
		CKB(cudaMemcpy(h->d_demod, st.data(), st.size() * sizeof(demod_state), cudaMemcpyHostToDevice));
	}
#undef CKB
	*out = h;
	return SONDE_OK;
}

void sonde_b200_destroy(sonde_b200 *h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	if (h->rate_registered) rate_registry(h->device, h->cfg.samplerate, -1);
	if (h->stream) cudaStreamSynchronize(h->stream);
	cudaFree(h->d_group_chan); cudaFree(h->d_group_type); cudaFree(h->d_types);
	cudaFree(h->d_demod); cudaFree(h->d_afsk); cudaFree(h->d_framer); cudaFree(h->d_ring);
	cudaFree(h->d_prof); cudaFree(h->d_in_row); cudaFree(h->d_active); cudaFree(h->d_cls_rows); cudaFree(h->d_cls_mask);
	for (int b = 0; b < 2; b++) {
		cudaFree(h->d_recs[b]); cudaFree(h->d_counts[b]); cudaFree(h->d_in[b]); cudaFree(h->d_in16[b]); cudaFree(h->d_nbits[b]);
		if (h->ev_demod[b]) cudaEventDestroy(h->ev_demod[b]);
		if (h->evf[b]) cudaEventDestroy(h->evf[b]);
		if (h->ev_copied[b]) cudaEventDestroy(h->ev_copied[b]);
		if (h->ev_done[b]) cudaEventDestroy(h->ev_done[b]);
	}
	cudaFree(h->d_soft);
	if (h->cstream) cudaStreamDestroy(h->cstream);
	if (h->dstream) cudaStreamDestroy(h->dstream);
	if (h->fstream) cudaStreamDestroy(h->fstream);
	if (h->ev_fork) cudaEventDestroy(h->ev_fork);
	for (int i = 0; i < 3; i++) {
		if (h->pstream[i]) cudaStreamDestroy(h->pstream[i]);
		if (h->ev_piece[i]) cudaEventDestroy(h->ev_piece[i]);
	}
	for (int v = 0; v < 4; v++) {
		if (h->vstream[v]) cudaStreamDestroy(h->vstream[v]);
		if (h->ev_join[v]) cudaEventDestroy(h->ev_join[v]);
	}
	if (h->h_counts) cudaFreeHost(h->h_counts);
	for (auto &e : h->ev)
		if (e) cudaEventDestroy(e);
	if (h->stream) cudaStreamDestroy(h->stream);
	delete h;
}

/* which virtual channels run: a fixed channel always; an AUTO channel's locked decoder once it has locked,
 * otherwise the decoders its `plausible` mask allows (all seven unless the pre-classifier narrowed it) */
static void compute_active(const sonde_b200 *h, std::vector<int32_t> &active)
{
	active.assign(h->cfg.n_channels, 1);
	for (int c = 0; c < h->n_user; c++) {
		if (h->user_types[c] != SONDE_AUTO) continue;
		for (int k = 0; k < SONDE_NTYPES; k++) {
			const int t = h->types[h->slot0[c] + k];
			const bool on = h->locked[c] != SONDE_AUTO ? (t == h->locked[c]) : ((h->plausible[c] >> t) & 1u);
			active[h->slot0[c] + k] = on ? 1 : 0;
		}
	}
}

/* regroup the active virtual channels (same kernels, fewer CTAs) and publish the new tables; the tables were sized
 * at create for one group per channel, so rebuilt ones always fit */
static int apply_active(sonde_b200 *h)
{
	std::vector<int32_t> active, gchan, gtype;
	compute_active(h, active);
	CK(cudaStreamSynchronize(h->stream));        /* launches in flight still read the old tables and group counts */
	CK(cudaStreamSynchronize(h->fstream));       /* ... and a framer kernel on its own stream reads d_active */
	build_groups(h, &active, gchan, gtype);
	h->gchan_host = gchan;
	if (!gchan.empty()) {
		CK(cudaMemcpyAsync(h->d_group_chan, gchan.data(), gchan.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
		CK(cudaMemcpyAsync(h->d_group_type, gtype.data(), gtype.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
	}
	CK(cudaMemcpyAsync(h->d_active, active.data(), active.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
	CK(cudaStreamSynchronize(h->stream));        /* the host vectors above go out of scope */
	return SONDE_OK;
}

/* AUTO pre-classifier policy, run at the start of a process call (before the demodulators are launched):
 *   - unlocked AUTO channels that have not been looked at yet are classified on this buffer; a channel whose
 *     run-length histogram is unambiguous only keeps the decoders of that family, the others stay "try all";
 *   - a channel narrowed more than FALLBACK_SECONDS of signal ago that still has not locked gets all seven back
 *     (the reference's behaviour), so a wrong guess can only delay the lock. */
static int auto_preclassify(sonde_b200 *h, const void *d_in, size_t len, size_t row_stride, int is_iq)
{
	const double FALLBACK_SECONDS = 3.0;
	bool changed = false;
	std::vector<int32_t> rows, users;
	for (int c = 0; c < h->n_user; c++) {
		if (h->user_types[c] != SONDE_AUTO || h->locked[c] != SONDE_AUTO) continue;
		if (h->narrowed_at[c] >= 0 &&
		    (double)(h->samples_seen - h->narrowed_at[c]) > FALLBACK_SECONDS * h->cfg.samplerate) {
			h->plausible[c] = (1u << SONDE_NTYPES) - 1u;
			h->narrowed_at[c] = -2;                          /* fell back: never narrowed again */
			changed = true;
		} else if (h->narrowed_at[c] == -1 && h->classify_pending) {
			rows.push_back(c);                               /* input row of user channel c is c */
			users.push_back(c);
		}
	}
	if (!rows.empty()) {
		const int n = (int)rows.size();
		if (!h->d_cls_rows) {
			CK(cudaMalloc(&h->d_cls_rows, (size_t)h->n_user * sizeof(int32_t)));
			CK(cudaMalloc(&h->d_cls_mask, (size_t)h->n_user * sizeof(uint32_t)));
		}
		std::vector<uint32_t> mask(n);
		CK(cudaMemcpyAsync(h->d_cls_rows, rows.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
		CK(sonde_launch_auto_classify(d_in, row_stride, (int)len, is_iq, h->d_cls_rows, n, h->d_cls_mask, h->stream));
		h->launches++;
		CK(cudaMemcpyAsync(mask.data(), h->d_cls_mask, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
		CK(cudaStreamSynchronize(h->stream));
		const uint32_t all = (1u << SONDE_NTYPES) - 1u;
		for (int i = 0; i < n; i++) {
			const int c = users[i];
			if (mask[i] != all && mask[i] != 0) {
				h->plausible[c] = mask[i];
				h->narrowed_at[c] = h->samples_seen;
				changed = true;
			} else if ((int)len >= 2048) {
				h->narrowed_at[c] = -2;                      /* looked at, ambiguous: stays try-all */
			}
		}
		h->classify_pending = false;
		for (int c = 0; c < h->n_user; c++)
			if (h->user_types[c] == SONDE_AUTO && h->locked[c] == SONDE_AUTO && h->narrowed_at[c] == -1) h->classify_pending = true;
	}
	return changed ? apply_active(h) : SONDE_OK;
}

static int run_chunk(sonde_b200 *h, const void *d_in, size_t len, size_t row_stride, int is_iq)
{
	if (!h) return SONDE_ERR_ARG;
	if (!d_in || len == 0) return fail(h, SONDE_ERR_ARG, "null input or zero length");
	if (len > (size_t)h->cfg.max_chunk_len) return fail(h, SONDE_ERR_TOOLONG, "len > max_chunk_len");
	if (row_stride < len) return fail(h, SONDE_ERR_ARG, "row_stride < len");
	CK(cudaSetDevice(h->device));
	if (h->has_auto && (h->cfg.reserved & 8)) {
		const int rc = auto_preclassify(h, d_in, len, row_stride, is_iq);
		if (rc != SONDE_OK) return rc;
	}
	h->samples_seen += (int64_t)len;

	demod_params dp;
	memset(&dp, 0, sizeof(dp));
	dp.in = d_in;
	dp.row_stride = row_stride;
	dp.len = (int32_t)len;
	dp.is_iq = is_iq;
	{
		/* bulk (TMA) copies need 16-byte aligned row segments: base, row pitch and the last tile's length */
		const size_t esz = is_iq ? 8 : 4;
		dp.use_tma = ((uintptr_t)d_in % 16 == 0) && ((row_stride * esz) % 16 == 0) && ((len * esz) % 16 == 0) &&
		             !(h->cfg.reserved & 2);
	}
	dp.fm_gain = h->cfg.fm_gain;
	dp.n_groups = h->n_groups;
	dp.group_chan = h->d_group_chan;
	dp.group_type = h->d_group_type;
	dp.st = h->d_demod;
	dp.ast = h->d_afsk;
	dp.ring = h->d_ring;
	dp.ring_bytes = h->ring_bytes;
	dp.soft = h->d_soft;
	dp.soft_stride = h->soft_stride;
	dp.prof = h->d_prof;
	dp.in_row = h->d_in_row;
	/* K1 warp placement: TM = warp 0 and LD = warp 4 on SMSP0, AG = warp 1 on SMSP1, twelve parallel-work warps on
	 * SMSP2/3 (ids 2,3,6,7,...) plus, per kernel variant, the measured best number of extra ones beside the serial
	 * warps (profiles/README.md); SONDE_PW_MASK overrides it for experiments */
	static const uint32_t mask_env = [] {
		const char *e = getenv("SONDE_PW_MASK");
		return e ? (uint32_t)strtoul(e, nullptr, 16) : 0u;
	}();
	static const uint32_t kPwMask[3] = {0xCCCFCCu /* RS41 */, 0xCCDFCCu /* DFM, iMS-100, MRZ-N1: their timing lane has slack for a second PW warp on its SMSP */, 0xCCEECCu /* M10/M20: two PW warps beside the AGC warp */};
	const int par = (int)(h->n_issued & 1);
	dp.nbits_out = h->d_nbits[par];
	static const int tpc_env = getenv("SONDE_TPC_PAIRS") ? atoi(getenv("SONDE_TPC_PAIRS")) : -1;     /* experiment switch */
	/* the demodulator of call i+2 appends to ring positions the framer of call i may still be reading */
	if (h->n_issued >= 2) CK(cudaStreamWaitEvent(h->stream, h->ev_done[par], 0));
	CK(cudaEventRecord(h->ev[0], h->stream));
	/* production kernel: the warp-specialised pipeline; reserved bit 0 selects the phase-by-phase
	 * kernel of demod.cu (kept as an independent cross-check for the tests).  A batch with several kernel
	 * variants (sonde mixes) forks them onto their own streams so that they share the GPU. */
	int n_variants = 0;
	for (int v = 0; v < 4; v++) n_variants += h->groups_v[v] > 0;
	static const bool no_fork = getenv("SONDE_NO_FORK") != nullptr;                /* experiment switch */
	const bool fork = n_variants > 1 && !no_fork;
	/* variants sharing the GPU: clusters of two, so that TPC siblings run the same code (pipe_common.cuh) */
	dp.tpc_pairs = tpc_env >= 0 ? tpc_env : (fork ? 1 : 0);
	if (fork) CK(cudaEventRecord(h->ev_fork, h->stream));
	int base = 0;
	for (int v = 0; v < 4; v++) {
		if (h->groups_v[v]) {
			cudaStream_t st = fork ? h->vstream[v] : h->stream;
			if (fork) CK(cudaStreamWaitEvent(st, h->ev_fork, 0));
			if (v == 3 && (h->cfg.reserved & 1)) CK(sonde_launch_demod_afsk(&dp, base, h->groups_v[v], st));
			else if (v == 3)               CK(sonde_launch_demod_pipe_afsk(&dp, base, h->groups_v[v], (h->cfg.reserved >> 2) & 1, st));
			else if (h->cfg.reserved & 1)  CK(sonde_launch_demod_gfsk(&dp, base, h->groups_v[v], v == 2 ? 2 : 1, st));
			else {
				dp.pw_mask = mask_env ? mask_env : kPwMask[v];
				static const bool no_2d = getenv("SONDE_NO_TMA2D") != nullptr;              /* experiment switch */
				dp.tma_box_rows = (dp.use_tma && !no_2d) ? h->box_rows_v[v] : 0;
				dp.tma_row_step = h->row_step_v[v];
				dp.n_rows = h->n_user;
				CK(sonde_launch_demod_pipe(&dp, base, h->groups_v[v], v, st));
			}
			h->launches++;
			if (fork) {
				CK(cudaEventRecord(h->ev_join[v], st));
				CK(cudaStreamWaitEvent(h->stream, h->ev_join[v], 0));
			}
		}
		base += h->groups_v[v];
	}
	CK(cudaEventRecord(h->ev[1], h->stream));
	CK(cudaEventRecord(h->ev_demod[par], h->stream));

	frame_params fp;
	memset(&fp, 0, sizeof(fp));
	fp.n_channels = h->cfg.n_channels;
	fp.types = h->d_types;
	fp.nbits = h->d_nbits[par];
	fp.fst = h->d_framer;
	fp.ring = h->d_ring;
	fp.ring_bytes = h->ring_bytes;
	fp.recs = h->d_recs[par];
	fp.max_frames = h->max_frames;
	fp.chunk_index = h->chunk_index;
	fp.counts = h->d_counts[par];
	fp.active = h->d_active;
	static const bool frame_serial = getenv("SONDE_FRAME_SERIAL") != nullptr;      /* experiment switch */
	static const int frame_skip = getenv("SONDE_FRAME_SKIP") ? atoi(getenv("SONDE_FRAME_SKIP")) : 0;
	fp.skip_warps = frame_skip;
	static const int frame_work = getenv("SONDE_FRAME_WORK") ? atoi(getenv("SONDE_FRAME_WORK")) : 0;
	static const int frame_persist = getenv("SONDE_FRAME_PERSIST") ? atoi(getenv("SONDE_FRAME_PERSIST")) : 0;
	fp.work_warps = frame_work;
	fp.persist = frame_persist;
	/* The framer runs in order behind the demodulator(s) of its call.  Measured with CUPTI timelines (tools/timeline.py,
	 * profiles/r2_step_experiments.md): on its own stream beside the next call's demodulator it costs that kernel — latency-
	 * bound serial warps, one CTA per SM — as much as it saves or more (config 2: 0.668 ms per step beside, 0.654 in order;
	 * config 5, forked variants: 2.16 against 1.74), however its warps are placed or throttled.  SONDE_FRAME_OVERLAP=1 restores
	 * the separate stream for experiments. */
	static const bool frame_overlap = getenv("SONDE_FRAME_OVERLAP") != nullptr;
	cudaStream_t fs_ = (frame_overlap && !frame_serial && !fork) ? h->fstream : h->stream;
	CK(cudaStreamWaitEvent(fs_, h->ev_demod[par], 0));
	CK(cudaEventRecord(h->evf[0], fs_));
	CK(sonde_launch_frames(&fp, fs_));
	h->launches++;
	CK(cudaEventRecord(h->evf[1], fs_));
	CK(cudaEventRecord(h->ev_done[par], fs_));
	h->have_timing = true;
	h->chunk_index++;
	h->n_issued++;
	return SONDE_OK;
}

int sonde_b200_process_iq_device(sonde_b200 *h, const void *d_iq, size_t len, size_t row_stride)
{
	return run_chunk(h, d_iq, len, row_stride, 1);
}

int sonde_b200_process_fm_device(sonde_b200 *h, const void *d_fm, size_t len, size_t row_stride)
{
	return run_chunk(h, d_fm, len, row_stride, 0);
}

static int process_host(sonde_b200 *h, const float *src, size_t len, int is_iq)
{
	if (!h) return SONDE_ERR_ARG;
	if (!src || len == 0) return fail(h, SONDE_ERR_ARG, "null input or zero length");
	if (len > (size_t)h->cfg.max_chunk_len) return fail(h, SONDE_ERR_TOOLONG, "len > max_chunk_len");
	CK(cudaSetDevice(h->device));
	const size_t esz = is_iq ? 2 * sizeof(float) : sizeof(float);
	const size_t need = (size_t)h->n_user * h->cfg.max_chunk_len * 2 * sizeof(float);
	const int par = (int)(h->n_issued & 1);
	if (!h->d_in[par]) CK(cudaMalloc(&h->d_in[par], need));
	/* copy on the copy stream once the kernels that last read this staging buffer are done; the compute
	 * stream then waits for the copy.  With pinned `src` the H2D of this call overlaps the previous call's kernels. */
	if (h->n_issued >= 2) CK(cudaStreamWaitEvent(h->cstream, h->ev_done[par], 0));
	CK(cudaMemcpyAsync(h->d_in[par], src, (size_t)h->n_user * len * esz, cudaMemcpyHostToDevice, h->cstream));
	CK(cudaEventRecord(h->ev_copied[par], h->cstream));
	CK(cudaStreamWaitEvent(h->stream, h->ev_copied[par], 0));
	return run_chunk(h, h->d_in[par], len, len, is_iq);
}

/* int16 IQ -> complex64, 4 samples (16 B in, 32 B out) per thread step; pure streaming, grid-stride */
__global__ void __launch_bounds__(256) s16_to_c64_kernel(const int4 *__restrict__ in, float4 *__restrict__ out, size_t n4,
                                                         const short2 *__restrict__ in_tail, float2 *__restrict__ out_tail,
                                                         int n_tail, float scale)
{
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
		const int4 q = __ldg(in + i);
		const short2 a = *reinterpret_cast<const short2 *>(&q.x), b = *reinterpret_cast<const short2 *>(&q.y);
		const short2 c = *reinterpret_cast<const short2 *>(&q.z), d = *reinterpret_cast<const short2 *>(&q.w);
		out[2 * i]     = make_float4(__fmul_rn((float)a.x, scale), __fmul_rn((float)a.y, scale),
		                             __fmul_rn((float)b.x, scale), __fmul_rn((float)b.y, scale));
		out[2 * i + 1] = make_float4(__fmul_rn((float)c.x, scale), __fmul_rn((float)c.y, scale),
		                             __fmul_rn((float)d.x, scale), __fmul_rn((float)d.y, scale));
	}
	if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) {
		const short2 a = in_tail[threadIdx.x];
		out_tail[threadIdx.x] = make_float2(__fmul_rn((float)a.x, scale), __fmul_rn((float)a.y, scale));
	}
}

int sonde_b200_process_iq_s16(sonde_b200 *h, const int16_t *iq, size_t len, float scale)
{
	if (!h) return SONDE_ERR_ARG;
	if (!iq || len == 0) return fail(h, SONDE_ERR_ARG, "null input or zero length");
	if (len > (size_t)h->cfg.max_chunk_len) return fail(h, SONDE_ERR_TOOLONG, "len > max_chunk_len");
	CK(cudaSetDevice(h->device));
	const size_t cap = (size_t)h->n_user * h->cfg.max_chunk_len;
	const int par = (int)(h->n_issued & 1);
	if (!h->d_in[par]) CK(cudaMalloc(&h->d_in[par], cap * 2 * sizeof(float)));
	if (!h->d_in16[par]) CK(cudaMalloc(&h->d_in16[par], cap * 2 * sizeof(int16_t)));
	const size_t n = (size_t)h->n_user * len;                    /* complex samples */
	if (h->n_issued >= 2) CK(cudaStreamWaitEvent(h->cstream, h->ev_done[par], 0));
	CK(cudaMemcpyAsync(h->d_in16[par], iq, n * 2 * sizeof(int16_t), cudaMemcpyHostToDevice, h->cstream));
	const size_t n4 = n / 4;
	const int n_tail = (int)(n % 4);
	const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? ((n4 + 255) / 256 ? (n4 + 255) / 256 : 1) : 148 * 8);
	s16_to_c64_kernel<<<blocks, 256, 0, h->cstream>>>(
		static_cast<const int4 *>(h->d_in16[par]), static_cast<float4 *>(h->d_in[par]), n4,
		static_cast<const short2 *>(h->d_in16[par]) + 4 * n4, static_cast<float2 *>(h->d_in[par]) + 4 * n4, n_tail, scale);
	CK(cudaGetLastError());
	h->launches++;
	CK(cudaEventRecord(h->ev_copied[par], h->cstream));
	CK(cudaStreamWaitEvent(h->stream, h->ev_copied[par], 0));
	return run_chunk(h, h->d_in[par], len, len, 1);
}

/* The channel block of this handle lives on ANOTHER GPU of the box (one front-end GPU feeding its peers, SURVEY.md
 * §8e): pull it over NVLink with the copy engine (cudaMemcpyPeerAsync on this handle's copy stream: no SMs, nothing
 * that competes with the demodulator for CTAs), then decode it.  Same double-buffered protocol as the host entry
 * points: the pull of call i+1 overlaps the kernels of call i. */
int sonde_b200_process_iq_peer(sonde_b200 *h, int src_device, const void *d_iq_src, size_t len, size_t row_stride)
{
	if (!h) return SONDE_ERR_ARG;
	if (!d_iq_src || len == 0) return fail(h, SONDE_ERR_ARG, "null input or zero length");
	if (len > (size_t)h->cfg.max_chunk_len) return fail(h, SONDE_ERR_TOOLONG, "len > max_chunk_len");
	if (row_stride < len) return fail(h, SONDE_ERR_ARG, "row_stride < len");
	CK(cudaSetDevice(h->device));
	if (src_device != h->device && !((h->peers_enabled >> src_device) & 1ull)) {
		int can = 0;
		CK(cudaDeviceCanAccessPeer(&can, h->device, src_device));
		if (can) {
			const cudaError_t e = cudaDeviceEnablePeerAccess(src_device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(h, SONDE_ERR_CUDA, "cudaDeviceEnablePeerAccess", e);
			(void)cudaGetLastError();
		}
		if (src_device >= 0 && src_device < 64) h->peers_enabled |= 1ull << src_device;      /* without P2P the copy is staged by the driver */
	}
	const size_t need = (size_t)h->n_user * h->cfg.max_chunk_len * 2 * sizeof(float);
	const int par = (int)(h->n_issued & 1);
	if (!h->d_in[par]) CK(cudaMalloc(&h->d_in[par], need));
	if (h->n_issued >= 2) CK(cudaStreamWaitEvent(h->cstream, h->ev_done[par], 0));
	const size_t esz = 2 * sizeof(float);
	if (row_stride == len) {
		/* one copy stream drives one copy engine, which moved ~245 GB/s of the link's ~750 GB/s on the B200 box; the
		 * block is therefore pulled as up to PEER_STREAMS pieces on as many streams */
		const size_t total = (size_t)h->n_user * len * esz;
		const int pieces = total >= ((size_t)8 << 20) ? SONDE_PEER_STREAMS : 1;
		const size_t step = ((total / pieces) + 255) & ~(size_t)255;
		if (pieces > 1) CK(cudaEventRecord(h->ev_fork, h->cstream));
		for (int i = 0; i < pieces; i++) {
			const size_t off = (size_t)i * step;
			if (off >= total) break;
			const size_t nbytes = std::min(step, total - off);
			cudaStream_t st = i == 0 ? h->cstream : h->pstream[i - 1];
			if (i > 0) {
				if (!st) {
					CK(cudaStreamCreateWithFlags(&h->pstream[i - 1], cudaStreamNonBlocking));
					CK(cudaEventCreateWithFlags(&h->ev_piece[i - 1], cudaEventDisableTiming));
					st = h->pstream[i - 1];
				}
				CK(cudaStreamWaitEvent(st, h->ev_fork, 0));
			}
			CK(cudaMemcpyPeerAsync(static_cast<char *>(h->d_in[par]) + off, h->device, static_cast<const char *>(d_iq_src) + off,
			                       src_device, nbytes, st));
			if (i > 0) {
				CK(cudaEventRecord(h->ev_piece[i - 1], st));
				CK(cudaStreamWaitEvent(h->cstream, h->ev_piece[i - 1], 0));
			}
		}
	} else {
		cudaMemcpy3DPeerParms pp;
		memset(&pp, 0, sizeof(pp));
		pp.srcDevice = src_device;
		pp.dstDevice = h->device;
		pp.srcPtr = make_cudaPitchedPtr(const_cast<void *>(d_iq_src), row_stride * esz, len * esz, (size_t)h->n_user);
		pp.dstPtr = make_cudaPitchedPtr(h->d_in[par], len * esz, len * esz, (size_t)h->n_user);
		pp.extent = make_cudaExtent(len * esz, (size_t)h->n_user, 1);
		CK(cudaMemcpy3DPeerAsync(&pp, h->cstream));
	}
	CK(cudaEventRecord(h->ev_copied[par], h->cstream));
	CK(cudaStreamWaitEvent(h->stream, h->ev_copied[par], 0));
	return run_chunk(h, h->d_in[par], len, len, 1);
}

int sonde_b200_process_iq(sonde_b200 *h, const float *iq, size_t len) { return process_host(h, iq, len, 1); }
int sonde_b200_process_fm(sonde_b200 *h, const float *fm, size_t len) { return process_host(h, fm, len, 0); }

int sonde_b200_max_frames(const sonde_b200 *h) { return h ? h->max_frames : SONDE_ERR_ARG; }
int sonde_b200_bits_stride(const sonde_b200 *h) { return h ? h->bits_stride : SONDE_ERR_ARG; }
int sonde_b200_soft_stride(const sonde_b200 *h) { return h ? h->soft_stride : SONDE_ERR_ARG; }
long sonde_b200_launch_count(const sonde_b200 *h) { return h ? h->launches : 0; }

int sonde_b200_sync(sonde_b200 *h)
{
	if (!h) return SONDE_ERR_ARG;
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->cstream));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaStreamSynchronize(h->fstream));
	CK(cudaStreamSynchronize(h->dstream));
	return SONDE_OK;
}

/* Make the main stream wait (on the device, not the host) for every framer kernel enqueued so far, so that an
 * event recorded on sonde_b200_stream() afterwards covers the whole work of the calls made up to now. */
int sonde_b200_join(sonde_b200 *h)
{
	if (!h) return SONDE_ERR_ARG;
	if (h->n_issued == 0) return SONDE_OK;
	CK(cudaSetDevice(h->device));
	CK(cudaStreamWaitEvent(h->stream, h->ev_done[(h->n_issued - 1) & 1], 0));
	return SONDE_OK;
}

/* Which call do fetch()/fetch_counts() serve: the oldest one not fetched yet among the last two issued
 * (results are double buffered; older unfetched results have been overwritten). */
static int fetch_slot(sonde_b200 *h, long *call)
{
	if (h->n_issued == 0) return -1;
	long c = h->n_fetched;
	if (c < h->n_issued - 2) c = h->n_issued - 2;
	if (c >= h->n_issued) c = h->n_issued - 1;          /* everything fetched already: serve the last call again */
	*call = c;
	return (int)(c & 1);
}

/* per-virtual-channel counters of one call into h->h_counts */
static int pull_counts(sonde_b200 *h, int par)
{
	const int V = h->cfg.n_channels;
	CK(cudaStreamWaitEvent(h->dstream, h->ev_done[par], 0));
	CK(cudaMemcpyAsync(h->h_counts, h->d_counts[par], (size_t)V * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->dstream));
	CK(cudaStreamSynchronize(h->dstream));
	return SONDE_OK;
}

/* AUTO policy (SD/decode.c:174-224): an undetermined channel locks to the first decoder, in the reference's
 * order, that produced a frame passing its gate in this call; the losers are switched off for later calls.
 *
 * Deliberate differences from the reference's loop, which decides inside a buffer, per decode() call:
 *   - granularity: the decision is taken per process call (buffer), at fetch time;
 *   - criterion: the reference locks when SondeData.fields != 0 — for RS41 any CRC-valid subframe, even when the
 *     Reed-Solomon decode of the frame failed — and, if several decoders report data in one iteration, the LAST
 *     set_active_decoder() call wins (decode.c:242-251).  Here the lock needs a frame that passes the decoder's FEC /
 *     checksum gate (rec.ok > 0: at least as strict) and the FIRST decoder in the reference's order wins.  On a signal
 *     that only one decoder can parse — every real transmission — both rules lock to the same decoder; the stricter gate
 *     can delay the lock of a heavily corrupted RS41 stream by the frames whose RS decode fails. */
static int auto_update(sonde_b200 *h)
{
	bool changed = false;
	for (int c = 0; c < h->n_user; c++) {
		if (h->user_types[c] != SONDE_AUTO || h->locked[c] != SONDE_AUTO) continue;
		const int v0 = h->slot0[c];
		for (int k = 0; k < SONDE_NTYPES; k++) {
			if (h->h_counts[2 * (v0 + k) + 1] > 0) {
				h->locked[c] = h->types[v0 + k];
				changed = true;
				break;
			}
		}
	}
	return changed ? apply_active(h) : SONDE_OK;
}

/* virtual channel whose results user channel c reports, -1 while an AUTO channel is undetermined */
static int report_slot(const sonde_b200 *h, int c)
{
	if (h->user_types[c] != SONDE_AUTO) return h->slot0[c];
	if (h->locked[c] == SONDE_AUTO) return -1;
	for (int k = 0; k < SONDE_NTYPES; k++)
		if (h->types[h->slot0[c] + k] == h->locked[c]) return h->slot0[c] + k;
	return -1;
}

static int fetch_counts_of(sonde_b200 *h, int par, int32_t *frames, int32_t *ok)
{
	int rc = pull_counts(h, par);
	if (rc) return rc;
	if (h->has_auto && (rc = auto_update(h))) return rc;
	for (int c = 0; c < h->n_user; c++) {
		const int v = report_slot(h, c);
		if (frames) frames[c] = v < 0 ? 0 : h->h_counts[2 * v];
		if (ok) ok[c] = v < 0 ? 0 : h->h_counts[2 * v + 1];
	}
	return SONDE_OK;
}

int sonde_b200_auto_plausible(sonde_b200 *h, uint32_t *masks)
{
	if (!h || !masks) return SONDE_ERR_ARG;
	for (int c = 0; c < h->n_user; c++) {
		if (h->user_types[c] != SONDE_AUTO) masks[c] = 1u << h->user_types[c];
		else if (h->locked[c] != SONDE_AUTO) masks[c] = 1u << h->locked[c];
		else masks[c] = h->plausible[c];
	}
	return SONDE_OK;
}

int sonde_b200_fetch_counts(sonde_b200 *h, int32_t *frames, int32_t *ok)
{
	if (!h) return SONDE_ERR_ARG;
	CK(cudaSetDevice(h->device));
	long call = 0;
	const int par = fetch_slot(h, &call);
	if (par < 0) return fail(h, SONDE_ERR_STATE, "no process call yet");
	const int rc = fetch_counts_of(h, par, frames, ok);
	if (rc == SONDE_OK && call >= h->n_fetched) h->n_fetched = call + 1;
	return rc;
}

int sonde_b200_detected_types(sonde_b200 *h, int32_t *types)
{
	if (!h || !types) return SONDE_ERR_ARG;
	for (int c = 0; c < h->n_user; c++) types[c] = h->locked[c];
	return SONDE_OK;
}

int sonde_b200_fetch_totals(sonde_b200 *h, int64_t *frames, int64_t *ok, int64_t *bits)
{
	if (!h) return SONDE_ERR_ARG;
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaStreamSynchronize(h->fstream));
	const int V = h->cfg.n_channels;
	std::vector<framer_state> fs(V);
	std::vector<demod_state> ds(V);
	CK(cudaMemcpy(fs.data(), h->d_framer, fs.size() * sizeof(framer_state), cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(ds.data(), h->d_demod, ds.size() * sizeof(demod_state), cudaMemcpyDeviceToHost));
	for (int c = 0; c < h->n_user; c++) {
		const int v = report_slot(h, c);
		if (frames) frames[c] = v < 0 ? 0 : fs[v].frames_total;
		if (ok) ok[c] = v < 0 ? 0 : fs[v].ok_total;
		if (bits) bits[c] = v < 0 ? 0 : (int64_t)ds[v].nbits;
	}
	return SONDE_OK;
}

int sonde_b200_fetch(sonde_b200 *h, sonde_frame_rec *recs, int32_t *counts)
{
	if (!h || !recs || !counts) return SONDE_ERR_ARG;
	CK(cudaSetDevice(h->device));
	long call = 0;
	const int par = fetch_slot(h, &call);
	if (par < 0) return fail(h, SONDE_ERR_STATE, "no process call yet");
	int rc = fetch_counts_of(h, par, counts, nullptr);
	if (rc) return rc;
	const int V = h->cfg.n_channels;
	/* The framer walks every window but only the first max_frames records of a channel are stored (frame.cu).  A channel
	 * that produced more (cannot happen for a well-formed stream, see max_frames) is reported truncated: one channel's
	 * garbage must not fail the batch. */
	for (int c = 0; c < h->n_user; c++)
		if (counts[c] > h->max_frames) {
			counts[c] = h->max_frames;
			h->n_truncated++;
			h->err = "frame records of a channel truncated to max_frames";
		}
	if (!h->has_auto) {
		CK(cudaMemcpyAsync(recs, h->d_recs[par], (size_t)V * h->max_frames * sizeof(sonde_frame_rec),
		                   cudaMemcpyDeviceToHost, h->dstream));
		CK(cudaStreamSynchronize(h->dstream));
	} else {
		/* only the reporting slot of each user channel crosses PCIe */
		for (int c = 0; c < h->n_user; c++) {
			const int v = report_slot(h, c);
			sonde_frame_rec *dst = recs + (size_t)c * h->max_frames;
			if (v < 0 || counts[c] == 0) continue;
			CK(cudaMemcpyAsync(dst, h->d_recs[par] + (size_t)v * h->max_frames, (size_t)counts[c] * sizeof(sonde_frame_rec),
			                   cudaMemcpyDeviceToHost, h->dstream));
		}
		CK(cudaStreamSynchronize(h->dstream));
	}
	if (call >= h->n_fetched) h->n_fetched = call + 1;
	return SONDE_OK;
}

int sonde_b200_fetch_bits(sonde_b200 *h, uint8_t *bits, int32_t *nbits)
{
	if (!h || !bits || !nbits) return SONDE_ERR_ARG;
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaStreamSynchronize(h->fstream));
	if (h->has_auto) return fail(h, SONDE_ERR_STATE, "parity taps are per decoder: not available with AUTO channels");
	const int C = h->cfg.n_channels;
	std::vector<demod_state> st(C);
	std::vector<uint8_t> ring((size_t)C * h->ring_bytes);
	CK(cudaMemcpy(st.data(), h->d_demod, st.size() * sizeof(demod_state), cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(ring.data(), h->d_ring, ring.size(), cudaMemcpyDeviceToHost));
	memset(bits, 0, (size_t)C * h->bits_stride);
	for (int c = 0; c < C; c++) {
		/* the bits of the last call are the last nsoft positions of the stream */
		const uint64_t end = st[c].nbits, n = (uint64_t)st[c].nsoft, start = end - n;
		if (n > (uint64_t)(h->bits_stride - 1) * 8) return fail(h, SONDE_ERR_STATE, "bit tap overflow");
		const uint8_t *r = ring.data() + (size_t)c * h->ring_bytes;
		uint8_t *o = bits + (size_t)c * h->bits_stride;
		for (uint64_t i = 0; i < n; i++) {
			const uint64_t p = start + i;
			const int b = (r[(p >> 3) & (h->ring_bytes - 1)] >> (7 - (p & 7))) & 1;
			o[i >> 3] |= (uint8_t)(b << (7 - (i & 7)));
		}
		nbits[c] = (int32_t)n;
	}
	return SONDE_OK;
}

int sonde_b200_fetch_soft(sonde_b200 *h, float *soft, int32_t *nsoft)
{
	if (!h || !soft || !nsoft) return SONDE_ERR_ARG;
	if (!h->d_soft) return fail(h, SONDE_ERR_STATE, "created without keep_soft");
	if (h->has_auto) return fail(h, SONDE_ERR_STATE, "parity taps are per decoder: not available with AUTO channels");
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaStreamSynchronize(h->fstream));
	const int C = h->cfg.n_channels;
	std::vector<demod_state> st(C);
	CK(cudaMemcpy(st.data(), h->d_demod, st.size() * sizeof(demod_state), cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(soft, h->d_soft, (size_t)C * h->soft_stride * sizeof(float), cudaMemcpyDeviceToHost));
	for (int c = 0; c < C; c++) nsoft[c] = st[c].nsoft;
	return SONDE_OK;
}

int sonde_b200_fetch_state(sonde_b200 *h, float *state /*[C][8]*/)
{
	if (!h || !state) return SONDE_ERR_ARG;
	if (h->has_auto) return fail(h, SONDE_ERR_STATE, "parity taps are per decoder: not available with AUTO channels");
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaStreamSynchronize(h->fstream));
	const int C = h->cfg.n_channels;
	std::vector<demod_state> st(C);
	CK(cudaMemcpy(st.data(), h->d_demod, st.size() * sizeof(demod_state), cudaMemcpyDeviceToHost));
	for (int c = 0; c < C; c++) {
		float *o = state + 8 * c;
		o[0] = st[c].agc_bias; o[1] = st[c].agc_avg; o[2] = st[c].t_phase; o[3] = st[c].t_freq;
		o[4] = st[c].t_prev; o[5] = (float)st[c].t_state; o[6] = st[c].disc_prev; o[7] = 0;
	}
	return SONDE_OK;
}

int sonde_b200_last_kernel_ms(sonde_b200 *h, float *demod_ms, float *frame_ms)
{
	if (!h) return SONDE_ERR_ARG;
	if (!h->have_timing) return fail(h, SONDE_ERR_STATE, "no process call yet");
	CK(cudaSetDevice(h->device));
	CK(cudaEventSynchronize(h->ev[1]));
	CK(cudaEventSynchronize(h->evf[1]));
	float a = 0, b = 0;
	CK(cudaEventElapsedTime(&a, h->ev[0], h->ev[1]));
	CK(cudaEventElapsedTime(&b, h->evf[0], h->evf[1]));
	if (demod_ms) *demod_ms = a;
	if (frame_ms) *frame_ms = b;
	return SONDE_OK;
}

int sonde_b200_modem_info(int type, int samplerate, float *taps, int taps_cap, float *consts)
{
	sonde_modem m;
	if (sonde_modem_init(&m, type, samplerate)) return SONDE_ERR_ARG;
	const int n = m.num_phases * SONDE_FIR_TAPS;
	for (int i = 0; taps && i < n && i < taps_cap; i++) taps[i] = m.taps[i];
	if (consts) {
		consts[0] = m.freq0; consts[1] = m.alpha; consts[2] = m.beta; consts[3] = m.max_fdev;
		consts[4] = (float)m.num_phases; consts[5] = (float)m.boxcar_len; consts[6] = m.f_mark; consts[7] = m.f_space;
	}
	return n;
}

void *sonde_b200_host_alloc(size_t bytes)
{
	void *p = nullptr;
	return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr;
}

void *sonde_b200_host_alloc_wc(size_t bytes)
{
	void *p = nullptr;
	return cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined) == cudaSuccess ? p : nullptr;
}

void sonde_b200_host_free(void *p)
{
	if (p) cudaFreeHost(p);
}

/* Diagnostics: enable (and read back) the pipeline kernel's per-CTA stall counters.
 * out[n_groups][16] cycles: [role*4 + {wait on input, wait on output slot, total}], roles PW, A1, A2, TM. */
int sonde_b200_debug_stalls(sonde_b200 *h, long long *out, int cap_groups)
{
	if (!h) return SONDE_ERR_ARG;
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->stream));
	if (!h->d_prof) {
		CK(cudaMalloc(&h->d_prof, (size_t)h->group_capacity * 16 * sizeof(long long)));
		CK(cudaMemset(h->d_prof, 0, (size_t)h->group_capacity * 16 * sizeof(long long)));
		return h->n_groups;
	}
	const int n = h->n_groups < cap_groups ? h->n_groups : cap_groups;
	if (out) CK(cudaMemcpy(out, h->d_prof, (size_t)n * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
	return n;
}

/* Diagnostics: the raw per-channel demodulator state (device_state.h demod_state, 256 bytes per virtual channel). */
int sonde_b200_debug_demod_state(sonde_b200 *h, void *out, size_t cap_bytes)
{
	if (!h || !out) return SONDE_ERR_ARG;
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->stream));
	const size_t need = (size_t)h->cfg.n_channels * sizeof(demod_state);
	if (cap_bytes < need) return fail(h, SONDE_ERR_ARG, "buffer too small");
	CK(cudaMemcpy(out, h->d_demod, need, cudaMemcpyDeviceToHost));
	return (int)sizeof(demod_state);
}

void *sonde_b200_stream(sonde_b200 *h) { return h ? (void *)h->stream : nullptr; }

}  /* extern "C" */
