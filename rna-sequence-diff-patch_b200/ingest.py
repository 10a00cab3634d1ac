"""ingest.py — sequence files -> packed database.

Mirrors the normalisation of the reference's importers (fa_import.py:61-62, import_xml.py:12-13:
T -> U, X -> N) and their record shape ({id: sequence}); replaces the MongoDB collection scan of
IRMethods.search_collection (IR:469) by a packed, device-resident database.  Unlike fa_import.py
(which imports only the first 500 records and drops the file's last one, fa_import.py:22,43-55) every
record is kept."""
from __future__ import annotations

import xml.etree.ElementTree as ET
from collections import OrderedDict

from .encoding import PackedSeqs, pack


def normalise(seq: str) -> str:
    return seq.strip().upper().replace("T", "U").replace("X", "N")


def read_seqxml(path: str) -> "OrderedDict[str, str]":
    """import_xml.py:4-17 — <entry id=...><RNAseq>...</RNAseq></entry>."""
    out: "OrderedDict[str, str]" = OrderedDict()
    for entry in ET.parse(path).getroot().findall("entry"):
        out[entry.get("id")] = normalise(entry.find("RNAseq").text)
    return out


def read_fasta(path: str) -> "OrderedDict[str, str]":
    """fa_import.py:39-62 — '>' header line, sequence possibly over several lines; all records."""
    out: "OrderedDict[str, str]" = OrderedDict()
    name, parts = None, []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if name is not None:
                    out[name] = normalise("".join(parts))
                name, parts = line[1:].split()[0] if len(line) > 1 else str(len(out)), []
            else:
                parts.append(line)
    if name is not None:
        out[name] = normalise("".join(parts))
    return out


class SequenceDB:
    """ids + packed sequences; `collection.find({})`-compatible so it can stand in for the Mongo
    collection the reference's callers pass around (IR:469)."""

    def __init__(self, records: "OrderedDict[str, str] | dict"):
        self.ids = list(records.keys())
        self.sequences = [records[k] for k in self.ids]
        self.packed: PackedSeqs = pack(self.sequences, bits=4)

    @classmethod
    def from_file(cls, path: str) -> "SequenceDB":
        return cls(read_seqxml(path) if path.lower().endswith(".xml") else read_fasta(path))

    def find(self, flt=None):
        return ({"_id": i, "sequence": s} for i, s in zip(self.ids, self.sequences))

    def __len__(self):
        return len(self.ids)
