/*
 * modem_tables.cpp — host-side construction of the per-sonde demodulator constants.
 *
 * The FIR taps and loop gains are computed ONCE per handle on the host and uploaded;
 * the kernels never call libm.  To get bit-identical taps the expressions below keep
 * the reference's evaluation types (which sub-expressions are double, which are float,
 * where the result is rounded to float) — see demod/dsp/filter.c:67-93 (tap formula:
 * raised cosine alpha = 0.99 times a Blackman window times 2/5*order/osf/10),
 * demod/dsp/filter.c:10-32 (polyphase layout), demod/dsp/timing.c:14-25,79-87 (loop
 * constants), demod/gfsk.c:17-35 and demod/afsk.c:16-48 (how they are parameterised).
 * tests/test_tables.py checks every tap and constant against the reference itself.
 *
 * Compile with -ffp-contract=off (x86-64 baseline has no FMA, like the reference build).
 */
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "sonde_params.h"

namespace {

const double kPi = 3.14159265358979323846;   /* M_PI */

struct proto {
	int afsk, baud, frame_bits, sync_len, data_len;
	uint64_t syncword;
	double f_mark, f_space;
};

/* indexed by enum sonde_type (include/sonde_b200.h) */
const proto kProto[SONDE_NTYPES_] = {
	/* RS41   */ {0, 4800, 4144, 64, 518, 0x086d53884469481fULL, 0, 0},
	/* DFM09  */ {0, 2500,  560, 32,  35, 0x9a995a55ULL,         0, 0},
	/* M10    */ {0, 9600, 1664, 48, 104, 0x66666666b366ULL,     0, 0},
	/* IMS100 */ {0, 2400, 1200, 48,  75, 0xaaa56a659a99ULL,     0, 0},
	/* MRZN1  */ {0, 2400,  816, 64,  51, 0x666666666555a599ULL, 0, 0},
	/* IMET4  */ {1, 1200,  600, 16,  60, 0xFF40ULL,             2200.0, 1200.0},
	/* C50    */ {1, 2380,   90, 20,   9, 0x005FFULL,            4700.0, 2900.0},
};

/* One tap of the windowed raised-cosine low-pass.  `k` indexes the full (taps*osf)-long
 * prototype, `ntaps` is its length. */
float lowpass_tap(float cutoff, int k, unsigned ntaps, float osf, float rolloff)
{
	const int order = (int)((ntaps - 1) / 2);
	const float norm = (float)(2.0 / 5.0 * order / osf / 10.0);
	const float t = (float)std::abs(order - k) / osf;
	float rc;

	if (t == 0) {
		rc = cutoff;
	} else if (2 * rolloff * t * cutoff == 1) {
		rc = (float)(-kPi / (4 * cutoff) * sinf((float)(kPi / (2 * rolloff))) / (kPi / (2 * rolloff)));
	} else {
		rc = (float)(sinf((float)(kPi * t * cutoff)) / (kPi * t)
		             * cosf((float)(kPi * rolloff * t * cutoff))
		             / (1 - powf(2 * rolloff * t * cutoff, 2)));
	}

	const float window = (float)(0.42
	                             - 0.5 * cosf((float)(2 * kPi * k / (ntaps - 1)))
	                             + 0.08 * cosf((float)(4 * kPi * k / (ntaps - 1))));
	return norm * rc * window;
}

}  // namespace

extern "C" int sonde_modem_init(sonde_modem *m, int type, int samplerate)
{
	if (!m || type < 0 || type >= SONDE_NTYPES_ || samplerate <= 0) return -1;
	const proto &p = kProto[type];

	memset(m, 0, sizeof(*m));
	m->type = type;
	m->afsk = p.afsk;
	m->baud = p.baud;
	m->frame_bits = p.frame_bits;
	m->sync_len = p.sync_len;
	m->data_len = p.data_len;
	m->syncword = p.syncword;

	const float sym_freq = (float)p.baud / samplerate;
	const int phases = (int)(1 + (8 * sym_freq));
	if (phases < 1 || phases > SONDE_MAX_PHASES) return -1;
	m->num_phases = phases;

	/* low-pass: cutoff 3 x symbol rate, split over the polyphase branches */
	float cutoff = 3 * sym_freq;
	cutoff /= phases;
	for (int ph = 0; ph < phases; ph++)
		for (int i = 0; i < SONDE_FIR_TAPS; i++)
			m->taps[ph * SONDE_FIR_TAPS + i] =
				lowpass_tap(cutoff, i * phases + ph, (unsigned)(SONDE_FIR_TAPS * phases), (float)phases, (float)0.99);

	/* Gardner loop: NCO counts 2 per symbol; 2nd-order loop, zeta 0.707, bw = f_sym/100 */
	const float f = sym_freq / phases;
	const float zeta = (float)0.707;
	const float bw = sym_freq / phases / 100;
	m->freq0 = 2 * f;
	m->max_fdev = f / (1 << 8);
	const float denom = (1 + 2 * zeta * bw + bw * bw);
	m->alpha = 4 * zeta * bw / denom;
	m->beta = 4 * bw * bw / denom;

	if (p.afsk) {
		m->f_mark = (float)(2 * kPi * p.f_mark / samplerate);
		m->f_space = (float)(2 * kPi * p.f_space / samplerate);
		m->boxcar_len = (int)(size_t)(1.0 / sym_freq);
		if (m->boxcar_len < 1 || m->boxcar_len > SONDE_AFSK_MAXLEN) return -1;
	}
	return 0;
}
