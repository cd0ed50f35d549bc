"""Shared helpers for the stream-level parity tests."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STREAM_DIR = os.path.join(HERE, "golden", "streams")
SHA1 = json.load(open(os.path.join(HERE, "golden", "stream_sha1.json")))
LIB_DIR = os.path.join(ROOT, "losslessh264_b200", "_lib")

# Streams of the parity corpus that the host decoder does not cover: none any more (the damaged ones - BA_MW_D_IDR_LOST,
# Error_I_P - follow the reference's freeze-output behaviour, host/decoder.cpp `frozen`).  Kept as the place to list, with
# the reason, anything a later change has to exclude: never skip silently.
NOT_YET = {}
# CABAC streams decode, compress and decompress like the CAVLC ones (host CABAC parse; the model hand-off follows the reference's
# one-macroblock skip runs and end_of_slice stop bits; the decompress core has a CABAC slice-data writer).  The set is only
# needed where a test is about CAVLC itself (the host CAVLC writer check).
CABAC = {"QCIF_2P_I_allIPCM.264", "test_qcif_cabac.264", "tibbycabac.264", "syn_cif_cabac_4slices_ltr_6f.264", "syn_qcif_cabac_qp38_8f.264",
         "test_cif_I_CABAC_slice.264", "test_cif_P_CABAC_slice.264", "sl_cabac_sps_pps.264", "ec_cif_cabac_ltr_tail_lost.264"}


# Streams the reference itself refuses: it hands out 0 frames and a `.pip` that holds only the default stream.
# test_scalinglist_jm: every PPS carries weighted_bipred_idc != 0 (B slices), which ParsePps rejects (au_parser.cpp), so no slice
# finds its parameter set.  The parity requirement is "refused, 0 frames, default stream identical" - tested explicitly.
REFUSED = {"test_scalinglist_jm.264"}

# Streams of tools/make_fmo_stream.py (slice groups of map types 0 / 1, several slices per group, arbitrary slice order).  Decoding
# and compression are checked on all of them; the ones with I_PCM macroblocks cannot come back from a `.pip` (I_PCM samples are not
# in the model, as in the reference), and the decompress direction is not built for slice groups (the reference's aborts on them).
PCM_SYNTH = {"fmo_dispersed2_idc1.264", "fmo_interleaved3_2slices.264", "fmo_dispersed4_3slices_aso.264", "aso_5slices.264",
             "fmo_dispersed7_2slices_aso_qvga.264"}
FMO = {"fmo_dispersed2_idc1.264", "fmo_interleaved3_2slices.264", "fmo_dispersed4_3slices_aso.264", "fmo_dispersed7_2slices_aso_qvga.264",
       "fmo_dispersed3_2slices_aso_nopcm.264"}

# Every stream with a golden SHA1 MUST be present: a missing file is an error, not a smaller test set (round-1 verdict: three
# corpus streams were absent and dropped out of the parametrised lists without a trace).
_missing = sorted(f for f in SHA1 if not os.path.exists(os.path.join(STREAM_DIR, f)))
if _missing:
    raise RuntimeError(f"tests/golden/streams lacks {_missing}: stream_sha1.json names them, the parity corpus is incomplete")
_unpinned = sorted(f for f in os.listdir(STREAM_DIR) if f not in SHA1)
if _unpinned:
    raise RuntimeError(f"tests/golden/streams holds {_unpinned} without a golden SHA1 in stream_sha1.json")


def decodable_streams():
    return sorted(f for f in SHA1 if f not in NOT_YET and f not in REFUSED)


def cavlc_streams():
    return [f for f in decodable_streams() if f not in CABAC]
