// model_decode.cuh — decompress direction of the recompression model on sm_100a: k_model_decode + lh264_model_decode().
//
// One warp per stream.  Lane 0 walks model_decode_core.inl (bool decoder -> adaptive priors -> macroblock syntax -> CAVLC
// bits); the whole warp does the per-slice FreqImage::updateFrame pass (decoded_macroblock.h:120-165) over the other frame's
// context records.  The decoder is a chain of dependent loads (prior -> bit -> next prior), i.e. latency-bound per stream:
// throughput comes from the number of streams in flight (one CTA of one warp each, up to 32 CTAs per SM), the stream's
// bool-decoder states and scratch levels sit in shared memory, and the prior table, context frames and side streams stay in
// HBM/L2 (the hot priors of a stream are a few hundred cache lines).
#pragma once
#define LH264_CAVLC_TABLE __device__ const
#include "../host/cavlc_tables.h"
#define LH264_CABAC_TABLE __device__ const
#include "../host/cabac_tables.h"
#define MD_FN __device__
#define MD_FN_NOINLINE __device__ __noinline__
#define MD_TABLE __device__ const
#define MD_CLZ(x) __clz ((int) (x))
#ifdef MD_INT_DIV
#define MD_DIV(n, d) ((n) / (d))
#else
// DynProb::updateProb divides 256 * (c0 + 1) < 2^18 by c0 + c1 + 2 <= 515 once per decision: multiply by ceil(2^32 / d) instead
// (exact: the error term stays below 2^-14 while the quotient's fraction is at most 1 - 1/515)
__constant__ uint32_t c_md_rcp[520] = {
  0u, 0u, 2147483648u, 1431655766u, 1073741824u, 858993460u, 715827883u, 613566757u, 536870912u, 477218589u,
  429496730u, 390451573u, 357913942u, 330382100u, 306783379u, 286331154u, 268435456u, 252645136u, 238609295u, 226050911u,
  214748365u, 204522253u, 195225787u, 186737709u, 178956971u, 171798692u, 165191050u, 159072863u, 153391690u, 148102321u,
  143165577u, 138547333u, 134217728u, 130150525u, 126322568u, 122713352u, 119304648u, 116080198u, 113025456u, 110127367u,
  107374183u, 104755300u, 102261127u, 99882961u, 97612894u, 95443718u, 93368855u, 91382283u, 89478486u, 87652394u,
  85899346u, 84215046u, 82595525u, 81037119u, 79536432u, 78090315u, 76695845u, 75350304u, 74051161u, 72796056u,
  71582789u, 70409300u, 69273667u, 68174085u, 67108864u, 66076420u, 65075263u, 64103990u, 63161284u, 62245903u,
  61356676u, 60492498u, 59652324u, 58835169u, 58040099u, 57266231u, 56512728u, 55778797u, 55063684u, 54366675u,
  53687092u, 53024288u, 52377650u, 51746594u, 51130564u, 50529028u, 49941481u, 49367441u, 48806447u, 48258060u,
  47721859u, 47197443u, 46684428u, 46182445u, 45691142u, 45210183u, 44739243u, 44278014u, 43826197u, 43383509u,
  42949673u, 42524429u, 42107523u, 41698712u, 41297763u, 40904451u, 40518560u, 40139882u, 39768216u, 39403370u,
  39045158u, 38693400u, 38347923u, 38008561u, 37675152u, 37347542u, 37025581u, 36709123u, 36398028u, 36092163u,
  35791395u, 35495598u, 35204650u, 34918434u, 34636834u, 34359739u, 34087043u, 33818641u, 33554432u, 33294321u,
  33038210u, 32786010u, 32537632u, 32292988u, 32051995u, 31814573u, 31580642u, 31350127u, 31122952u, 30899046u,
  30678338u, 30460761u, 30246249u, 30034737u, 29826162u, 29620465u, 29417585u, 29217465u, 29020050u, 28825284u,
  28633116u, 28443493u, 28256364u, 28071682u, 27889399u, 27709467u, 27531842u, 27356480u, 27183338u, 27012373u,
  26843546u, 26676816u, 26512144u, 26349493u, 26188825u, 26030105u, 25873297u, 25718368u, 25565282u, 25414008u,
  25264514u, 25116768u, 24970741u, 24826401u, 24683721u, 24542671u, 24403224u, 24265352u, 24129030u, 23994231u,
  23860930u, 23729102u, 23598722u, 23469767u, 23342214u, 23216040u, 23091223u, 22967740u, 22845571u, 22724695u,
  22605092u, 22486740u, 22369622u, 22253717u, 22139007u, 22025474u, 21913099u, 21801865u, 21691755u, 21582751u,
  21474837u, 21367997u, 21262215u, 21157475u, 21053762u, 20951060u, 20849356u, 20748635u, 20648882u, 20550083u,
  20452226u, 20355296u, 20259280u, 20164166u, 20069941u, 19976593u, 19884108u, 19792477u, 19701685u, 19611723u,
  19522579u, 19434242u, 19346700u, 19259944u, 19173962u, 19088744u, 19004281u, 18920561u, 18837576u, 18755316u,
  18673771u, 18592933u, 18512791u, 18433337u, 18354562u, 18276457u, 18199014u, 18122225u, 18046082u, 17970575u,
  17895698u, 17821442u, 17747799u, 17674763u, 17602325u, 17530479u, 17459217u, 17388532u, 17318417u, 17248865u,
  17179870u, 17111424u, 17043522u, 16976156u, 16909321u, 16843010u, 16777216u, 16711936u, 16647161u, 16582886u,
  16519105u, 16455814u, 16393005u, 16330675u, 16268816u, 16207424u, 16146494u, 16086020u, 16025998u, 15966422u,
  15907287u, 15848588u, 15790321u, 15732481u, 15675064u, 15618063u, 15561476u, 15505298u, 15449523u, 15394149u,
  15339169u, 15284582u, 15230381u, 15176563u, 15123125u, 15070061u, 15017369u, 14965043u, 14913081u, 14861479u,
  14810233u, 14759338u, 14708793u, 14658592u, 14608733u, 14559212u, 14510025u, 14461170u, 14412642u, 14364440u,
  14316558u, 14268995u, 14221747u, 14174810u, 14128182u, 14081860u, 14035841u, 13990122u, 13944700u, 13899571u,
  13854734u, 13810185u, 13765921u, 13721941u, 13678240u, 13634817u, 13591669u, 13548793u, 13506187u, 13463848u,
  13421773u, 13379961u, 13338408u, 13297113u, 13256072u, 13215284u, 13174747u, 13134457u, 13094413u, 13054612u,
  13015053u, 12975733u, 12936649u, 12897800u, 12859184u, 12820798u, 12782641u, 12744711u, 12707004u, 12669521u,
  12632257u, 12595213u, 12558384u, 12521771u, 12485371u, 12449181u, 12413201u, 12377428u, 12341861u, 12306497u,
  12271336u, 12236375u, 12201612u, 12167047u, 12132676u, 12098500u, 12064515u, 12030721u, 11997116u, 11963698u,
  11930465u, 11897417u, 11864551u, 11831866u, 11799361u, 11767034u, 11734884u, 11702909u, 11671107u, 11639478u,
  11608020u, 11576732u, 11545612u, 11514658u, 11483870u, 11453247u, 11422786u, 11392487u, 11362348u, 11332368u,
  11302546u, 11272881u, 11243370u, 11214014u, 11184811u, 11155760u, 11126859u, 11098107u, 11069504u, 11041048u,
  11012737u, 10984572u, 10956550u, 10928670u, 10900933u, 10873335u, 10845878u, 10818558u, 10791376u, 10764330u,
  10737419u, 10710642u, 10683999u, 10657488u, 10631108u, 10604858u, 10578738u, 10552746u, 10526881u, 10501143u,
  10475530u, 10450043u, 10424678u, 10399437u, 10374318u, 10349319u, 10324441u, 10299682u, 10275042u, 10250519u,
  10226113u, 10201823u, 10177648u, 10153587u, 10129640u, 10105806u, 10082083u, 10058472u, 10034971u, 10011579u,
  9988297u, 9965122u, 9942054u, 9919094u, 9896239u, 9873489u, 9850843u, 9828301u, 9805862u, 9783525u,
  9761290u, 9739155u, 9717121u, 9695186u, 9673350u, 9651612u, 9629972u, 9608428u, 9586981u, 9565629u,
  9544372u, 9523210u, 9502141u, 9481165u, 9460281u, 9439489u, 9418788u, 9398178u, 9377658u, 9357228u,
  9336886u, 9316632u, 9296467u, 9276388u, 9256396u, 9236489u, 9216669u, 9196933u, 9177281u, 9157713u,
  9138229u, 9118827u, 9099507u, 9080270u, 9061113u, 9042037u, 9023041u, 9004125u, 8985288u, 8966529u,
  8947849u, 8929246u, 8910721u, 8892272u, 8873900u, 8855603u, 8837382u, 8819235u, 8801163u, 8783165u,
  8765240u, 8747388u, 8729609u, 8711902u, 8694266u, 8676702u, 8659209u, 8641786u, 8624433u, 8607149u,
  8589935u, 8572790u, 8555712u, 8538703u, 8521761u, 8504886u, 8488078u, 8471336u, 8454661u, 8438050u,
  8421505u, 8405025u, 8388608u, 8372256u, 8355968u, 8339743u, 8323581u, 8307481u, 8291443u, 8275467u };
__device__ __forceinline__ uint32_t md_div (uint32_t num, uint32_t den) { return __umulhi (num, c_md_rcp[den]); }
#define MD_DIV(n, d) md_div (n, d)
#endif
// one stream per CTA: the bool-decoder states and TEST_PROB are CTA-wide shared variables, named directly so that the
// compiler emits LDS/STS instead of generic accesses through the context
#define MD_READER(c, tag) (lh264md::s_rd[tag])
#define MD_TEST(c) (lh264md::s_test)
#define MD_SHARED_STATE 1
#include "model_decode_core.inl"

// launch geometry: one warp per job
#ifndef MD_BLOCK
#define MD_BLOCK 32
#endif
#define MD_GRID(n_jobs) (n_jobs)
#ifndef MD_VARIANT_NAME
#define MD_VARIANT_NAME "default"
#endif

namespace lh264md {

static_assert (sizeof (MdSliceIn) == sizeof (lh264_decode_slice), "slice descriptor layout");
static_assert (sizeof (MdStream) % 8 == 0, "MdStream is copied by words");

struct MdJobDev {
  MdStream st;                         // initial state: pointers set by the host, everything else zero
  const MdSliceIn* slices;
  uint8_t* out; uint64_t cap;
  lh264_decode_slice_out* sout;
  const uint8_t* tag_buf[CT_COUNT];
  uint32_t tag_len[CT_COUNT];
  int32_t n_slices, error;
};

// FreqImage::updateFrame, warp-wide: isSkipped per record, and the length of every run of all-zero records written into the
// LAST `run` records ending at the non-zero one that closes it (the reference's off-by-one is kept; see k_model_frame).
__device__ void md_update_frame_warp (CtxRec* f, int len, int lane) {
  int run_start = 0;
  for (int base = 0; base < len; base += 32) {
    const int i = base + lane;
    bool zeroed = false;
    const bool valid = i < len;
    if (valid) {
      const uint4* q = reinterpret_cast<const uint4*> (&f[i]);
      const uint4 a = q[0], b = q[1];
      zeroed = ((a.z | a.w | b.x | b.y | b.z | b.w) == 0);
      uint8_t fl = (uint8_t) (a.x & 0xff);
      fl = zeroed ? (fl | 2) : (fl & ~2);
      f[i].flags = fl;
    }
    unsigned nzmask = __ballot_sync (0xffffffffu, valid && !zeroed);
    while (nzmask) {
      const int l = __ffs (nzmask) - 1;
      nzmask &= nzmask - 1;
      const int at = base + l, run = at - run_start;
      if (run > 0) for (int j = at - run + 1 + lane; j <= at; j += 32) f[j].cachedSkips = (uint16_t)run;
      run_start = at + 1;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__ (32, 28) k_model_decode (MdJobDev* jobs, int n_jobs) {   // 28 one-warp CTAs per SM: 4096 streams resident at once
  __shared__ MdStream st;
  __shared__ MdScratch sc;
  __shared__ CtxRec* s_frame;
  if ((int)blockIdx.x >= n_jobs) return;
  MdJobDev& job = jobs[blockIdx.x];
  const int lane = threadIdx.x;
  for (unsigned i = lane; i < sizeof (MdStream) / 4; i += 32) reinterpret_cast<uint32_t*> (&st)[i] = reinterpret_cast<const uint32_t*> (&job.st)[i];
  __syncwarp();
  md_scratch_init (sc, lane, 32);
  if (lane == 0) s_test = job.st.priors[P_TEST];
  if (lane < CT_COUNT) {
    md_reader_init (s_rd[lane], job.tag_buf[lane], job.tag_buf[lane] ? job.tag_len[lane] : 0u);
    if (job.tag_buf[lane]) atomicOr (&st.present, 1u << lane);
  }
  __syncwarp();
  MdCtx c;
  md_ctx_init (c, st.priors, st.rd);
  uint64_t off = 0;
  for (int s = 0; s < job.n_slices; s++) {
    MdSliceIn sl = job.slices[s];
    if (lane == 0) s_frame = md_slice_flip (st, sl.frame_num);
    __syncwarp();
    md_update_frame_warp (s_frame, st.frame_len, lane);
    __syncwarp();
    if (lane == 0) {
      const bool prior_valid = md_slice_resize (st, sl.mb_w, sl.mb_h);
      if (!st.error) {
        MdBits bw;
        mb_init (bw, job.out + off, job.cap - off);
        lh264_decode_slice_out so;
        so.reserved = 0; so.reserved2 = 0;
        so.n_mbs = (uint32_t)md_slice (st, c, sl, prior_valid, bw, sc, &so.pad_bits, &so.n_pad_bits);
        mb_flush (bw);
        so.byte_offset = off; so.n_bits = (uint32_t) (mb_bits (bw) - (uint64_t)bw.trim);
        job.sout[s] = so;
        off += (mb_bits (bw) + 7) / 8;
      }
    }
    __syncwarp();
    if (st.error) break;
  }
  if (lane == 0) job.error = st.error;
}

} // namespace lh264md
