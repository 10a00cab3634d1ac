"""Host-side 'next rows': ingest (FASTA / SeqXML -> packed DB) and edit-script wire formats. CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as G
    G.build()
    import rna_sequence_diff_patch_b200 as R
    from rna_sequence_diff_patch_b200 import eswire, ingest, sed
    return R, eswire, ingest, sed


def test_seqxml_and_fasta_ingest(pkg, golden, tmp_path):
    R, eswire, ingest, sed = pkg
    xml = tmp_path / "in.xml"
    dna = [s.replace("U", "T") for s in golden["xml_seqs"]]           # test_input.xml holds DNA letters
    xml.write_text('<?xml version="1.0"?>\n<seqXML>\n' + "\n".join(
        f'  <entry id="{i}" >\n    <RNAseq>{s}</RNAseq>\n  </entry>' for i, s in zip(golden["xml_ids"], dna)) + "\n</seqXML>\n")
    recs = ingest.read_seqxml(str(xml))
    assert [k for k, _ in recs] == golden["xml_ids"] and [v for _, v in recs] == golden["xml_seqs"]
    fa = tmp_path / "in.fa"
    # repeated titles and titles with a common first word: every record is a document of its own (fa_import.py:49),
    # the title is the whole header line (fa_import.py:56)
    fa.write_text("".join(f">{i} some description\n{s[:10]}\n{s[10:]}\n" for i, s in zip(golden["xml_ids"], dna)) +
                  ">last\nACGTX\n>dup x\nAAT\n>dup y\nCCX\n>dup x\nGG\n")
    recs = ingest.read_fasta(str(fa))
    assert [v for _, v in recs[:-4]] == golden["xml_seqs"]
    assert recs[-4:] == [("last", "ACGUN"), ("dup x", "AAU"), ("dup y", "CCN"), ("dup x", "GG")]
    assert recs[0][0] == golden["xml_ids"][0] + " some description"
    assert len(ingest.SequenceDB(recs)) == len(golden["xml_seqs"]) + 4
    db = ingest.SequenceDB.from_file(str(xml))
    assert len(db) == 25 and [d["sequence"] for d in db.find({})] == golden["xml_seqs"]
    from rna_sequence_diff_patch_b200.encoding import unpack
    codes, off = unpack(db.packed)
    assert O.decode(codes[off[3]:off[4]]) == golden["xml_seqs"][3]


def test_es_json_roundtrip_and_packed_conversions(pkg, golden, tmp_path):
    R, eswire, ingest, sed = pkg
    n = 0
    for c in golden["small"] + golden["medium"]:
        if "es" not in c:
            continue
        es = c["es"][0]
        path = tmp_path / "es.json"
        eswire.dump_es(es, str(path))
        assert path.read_text() == json.dumps({"edit_script": es}, indent=4)       # gui.py:636-639
        assert eswire.load_es(str(path)) == es
        op, oi, oj = eswire.es_to_packed(es)
        ops, ooi, ooj, _ = O.canonical_script(c["a"], c["b"], golden["user_costs" if c["user"] else "default_costs"])
        assert np.array_equal(op, ops) and np.array_equal(oi, ooi) and np.array_equal(oj, ooj)
        assert eswire.packed_to_es(op, oi, oj, c["a"], c["b"]) == es
        assert eswire.rev_es_from_packed(op, oi, oj, c["a"], c["b"]) == c["rev0"]  # == generate_rev_es(es)
        rop, roi, roj = eswire.rev_packed(op, oi, oj)                              # packed B -> A script
        code, back = O.patch_closed_codes(rop, roi, roj, O.encode(c["b"]), O.encode(c["a"]), O.encode(c["b"]))
        assert code == 0 and O.decode(back) == c["a"]
        n += 1
    assert n > 250


def test_pack_word_boundaries_threads_and_errors():
    """rsd_pack (host): every length around the 16- / 8-symbol word boundaries, both packings, the threaded path
    (>= 4 M symbols) and the error for a code that does not fit — in a full word and in a tail."""
    import numpy as np
    import rna_sequence_diff_patch_b200 as R
    from rna_sequence_diff_patch_b200.encoding import unpack
    rng = np.random.default_rng(5)
    for bits, alpha in ((2, 4), (4, 15)):
        lens = np.array(list(range(0, 70)) * 3, np.int64)
        off = np.zeros(lens.shape[0] + 1, np.int64); np.cumsum(lens, out=off[1:])
        codes = rng.integers(0, alpha, size=int(off[-1]), dtype=np.uint8)
        P = R.pack((codes, off), bits=bits)
        c2, o2 = unpack(P)
        assert np.array_equal(c2, codes) and np.array_equal(o2, off)
        per = 32 // bits
        assert np.array_equal(P.start, np.concatenate([[0], np.cumsum((lens + per - 1) // per)[:-1]]))
        want_mask = 0
        for v in np.unique(codes):
            want_mask |= 1 << int(v)
        assert P.symmask == want_mask
        for pos in (3, int(off[40]) + 35, int(off[-1]) - 1):            # first word, a tail, the very last symbol
            bad = codes.copy(); bad[pos] = 1 << bits
            try:
                R.pack((bad, off), bits=bits)
                assert False, "code out of range accepted"
            except R.RsdError as e:
                assert "does not fit" in str(e)
    # threaded path
    n = 60000
    lens = rng.integers(40, 120, size=n)
    off = np.zeros(n + 1, np.int64); np.cumsum(lens, out=off[1:])
    assert off[-1] >= (1 << 22)
    codes = rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)
    P = R.pack((codes, off))
    assert P.bits == 2 and P.symmask == 0xF
    sub = R.PackedSeqs(P.words, P.start[:2000], P.len[:2000], P.bits, P.symmask)
    assert np.array_equal(unpack(sub)[0], codes[:off[2000]])
    tail = R.PackedSeqs(P.words, P.start[-500:], P.len[-500:], P.bits, P.symmask)
    assert np.array_equal(unpack(tail)[0], codes[off[n - 500]:])
