"""ctypes mirror of include/longtr_b200_locus.h (``ltr_flat_locus`` / ``ltr_flat_read``).

A flat locus is the language-neutral form of what LongTR hands its hot path: the three
haplotype blocks wrapped by ``Haplotype`` (reference: src/SeqAlignment/Haplotype.h:34-50,
HaplotypeGenerator.cpp:580-607) and the pooled ``Alignment`` reads consumed by
``HapAligner::process_reads`` (src/SeqAlignment/HapAligner.h:137-138).
"""
import ctypes as C

DEFAULT_STUTTER = (0.95, 0.05, 0.05, 0.95, 0.01, 0.01)  # hipstr_main.cpp:362-363 fixed model


class FlatRead(C.Structure):
    _fields_ = [("start", C.c_int32), ("stop", C.c_int32), ("seq", C.c_char_p),
                ("qual", C.c_char_p), ("cigar", C.c_char_p)]


class FlatLocus(C.Structure):
    _fields_ = [("lflank", C.c_char_p),
                ("repeat_start", C.c_int32), ("repeat_end", C.c_int32),
                ("period", C.c_int32), ("n_alleles", C.c_int32),
                ("alleles", C.POINTER(C.c_char_p)),
                ("rflank", C.c_char_p),
                ("stutter", C.c_double * 6), ("motif", C.c_char_p),
                ("n_reads", C.c_int32), ("reads", C.POINTER(FlatRead)),
                ("indel_flank_len", C.c_int32), ("switch_old_align_len", C.c_int32),
                ("n_aln_params", C.c_int32), ("aln_params", C.c_float * 7),
                ("realign_to_hap", C.POINTER(C.c_uint8)), ("realign_read", C.POINTER(C.c_uint8))]


def _b(s):
    return s if isinstance(s, bytes) else s.encode()


def make_flat_locus(lflank, alleles, rflank, repeat_start, repeat_end, period, reads,
                    motif="A", stutter=DEFAULT_STUTTER, indel_flank_len=5,
                    switch_old_align_len=0, aln_params=None,
                    realign_to_hap=None, realign_read=None):
    """Build a FlatLocus. ``reads`` is a list of dicts/tuples (start, stop, seq, qual, cigar).

    Returns (locus, keepalive); keep ``keepalive`` referenced while the struct is in use.
    """
    keep = []
    L = FlatLocus()
    L.lflank = _b(lflank)
    L.rflank = _b(rflank)
    L.repeat_start, L.repeat_end, L.period = repeat_start, repeat_end, period
    arr = (C.c_char_p * len(alleles))(*[_b(a) for a in alleles])
    keep.append(arr)
    L.n_alleles = len(alleles)
    L.alleles = C.cast(arr, C.POINTER(C.c_char_p))
    L.stutter = (C.c_double * 6)(*stutter)
    L.motif = _b(motif)
    rarr = (FlatRead * max(1, len(reads)))()
    for i, r in enumerate(reads):
        if isinstance(r, dict):
            r = (r["start"], r["stop"], r["seq"], r["qual"], r["cigar"])
        rarr[i].start, rarr[i].stop = int(r[0]), int(r[1])
        rarr[i].seq, rarr[i].qual, rarr[i].cigar = _b(r[2]), _b(r[3]), _b(r[4])
    keep.append(rarr)
    L.n_reads = len(reads)
    L.reads = C.cast(rarr, C.POINTER(FlatRead))
    L.indel_flank_len = indel_flank_len
    L.switch_old_align_len = switch_old_align_len
    if aln_params is None:
        L.n_aln_params = 0
    else:
        assert len(aln_params) == 7
        L.n_aln_params = 7
        L.aln_params = (C.c_float * 7)(*aln_params)
    if realign_to_hap is not None:
        a = (C.c_uint8 * len(realign_to_hap))(*[1 if x else 0 for x in realign_to_hap])
        keep.append(a)
        L.realign_to_hap = C.cast(a, C.POINTER(C.c_uint8))
    if realign_read is not None:
        a = (C.c_uint8 * len(realign_read))(*[1 if x else 0 for x in realign_read])
        keep.append(a)
        L.realign_read = C.cast(a, C.POINTER(C.c_uint8))
    return L, keep
