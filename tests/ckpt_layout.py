"""Section map of a reference `.short` checkpoint file, used by the checkpoint tests to
compare two files field by field and to leave out the fields the reference never reads back before it
overwrites them (documented in gmix_b200/csrc/checkpoint.h). TEST INFRASTRUCTURE ONLY."""
import struct

CELLS, HORIZON, NIN, ROW, NOUT, HID = 50, 100, 307, 563, 256, 51

# scratch: written by the reference, rebuilt before any read after ReadCheckpoint
SCRATCH = {"ppmd.aux_unit", "ppmd.saved_pc", "ppmd.sq", "ppmd.sq_ptr", "ppmd.sqp_trf_trt", "stm.entropy", "stm.rotating_history",
           "lstm.nl0.error", "lstm.nl0.update", "lstm.nl0.transpose", "lstm.nl1.error", "lstm.nl1.update", "lstm.nl1.transpose",
           "lstm.nl2.error", "lstm.nl2.update", "lstm.nl2.transpose"}


def short_sections(blob):
    """Returns [(name, start, end)] covering the whole file."""
    out = []
    pos = 0

    def take(name, n):
        nonlocal pos
        out.append((name, pos, pos + n))
        pos += n

    take("basic.first_prediction", 1)
    take("ppmd.top_mid_bot", 12)
    take("ppmd.blist", 39 * 8)
    take("ppmd.glue", 8)
    take("ppmd.sa_size", 8)
    take("ppmd.ptext_unitsstart_lo_hi", 32)
    take("ppmd.aux_unit", 8)
    take("ppmd.found_state", 8)
    take("ppmd.max_context", 8)
    take("ppmd.saved_pc", 8)
    take("ppmd.orderfall_esccount", 8)
    take("ppmd.char_mask", 1024)
    take("ppmd.bsumm_rl_initrl_nummasked_prevsuccess", 20)
    take("ppmd.bin_summ", 25 * 64 * 2)
    take("ppmd.see2", 23 * 32 * 4 + 4)
    take("ppmd.sq", 1024 * 6)
    take("ppmd.sq_ptr", 4)
    take("ppmd.sqp_trf_trt", 256 * 12)
    take("ppmd.cxt_y", 8)
    nseq = struct.unpack_from("<i", blob, pos)[0]
    take("ppmd.heap_runs", 4 + 16 * nseq)
    runs = [struct.unpack_from("<QQ", blob, out[-1][1] + 4 + 16 * i) for i in range(nseq)]
    sa = struct.unpack_from("<Q", blob, [s for s in out if s[0] == "ppmd.sa_size"][0][1])[0]
    literal = sa - sum(c for c, _ in runs)
    take("ppmd.heap_bytes", literal)
    take("lstm.top_mid_bot", 12)
    take("lstm.probs", 1024)
    take("lstm.input_history", HORIZON * 4)
    take("lstm.hidden", HID * 4)
    take("lstm.hidden_error", CELLS * 4)
    take("lstm.layer_input", HORIZON * NIN * 4)
    take("lstm.output", HORIZON * NOUT * 4)
    take("lstm.epoch", 4)
    take("lstm.state_stateerr_storederr", 3 * CELLS * 4)
    take("lstm.tanh_ig_last", 3 * HORIZON * CELLS * 4)
    take("lstm.layer_epoch_updatesteps", 12)
    for g in range(3):
        take(f"lstm.nl{g}.error", CELLS * 4)
        take(f"lstm.nl{g}.ivar", HORIZON * 4)
        take(f"lstm.nl{g}.gamma_beta", 8 * CELLS * 4)
        take(f"lstm.nl{g}.state", HORIZON * CELLS * 4)
        take(f"lstm.nl{g}.update", CELLS * ROW * 4)
        take(f"lstm.nl{g}.m", CELLS * ROW * 4)
        take(f"lstm.nl{g}.v", CELLS * ROW * 4)
        take(f"lstm.nl{g}.transpose", HID * CELLS * 4)
        take(f"lstm.nl{g}.norm", HORIZON * CELLS * 4)
    take("match", 6 * 11)
    for k, log2 in enumerate([8, 8, 8, 16, 16, 16, 24, 24, 24]):
        n = struct.unpack_from("<I", blob, pos)[0]
        size = 1 << log2
        take(f"ih{k}", 4 + (8 * n if n < size // 2 else 4 * size) + 12)
    take("mixer", 33 * 24)
    take("stm.predictions", 90 * 4)
    take("stm.scalars_and_contexts", (2 + 7 + 9 + 9 + 15 + 2) * 4)
    take("stm.mixer_outputs", (24 + 8 + 1) * 4)
    take("stm.longest_match", 4)
    take("stm.bits_seen", 8)
    take("stm.entropy", 123 * 8)
    take("stm.lstm_prediction_context", 4)
    take("stm.rotating_history", 1000)
    take("stm.rotating_history_pos", 4)
    take("stm.recent_bytes", 40)
    assert pos == len(blob), (pos, len(blob))
    return out


def differing_sections(a, b):
    """Names of the sections in which two `.short` files of identical structure differ."""
    sa, sb = short_sections(a), short_sections(b)
    if [(n, e - s) for n, s, e in sa] != [(n, e - s) for n, s, e in sb]:
        return ["<structure>"]
    return [n for (n, s, e), (_, s2, e2) in zip(sa, sb) if a[s:e] != b[s2:e2]]
