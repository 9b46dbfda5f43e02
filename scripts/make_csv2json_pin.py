"""Generates tests/golden/csv2json_sample.out: the expected output of csv2json on the reference's
10-row sample (test/data/csv/csv_format1.sample.csv, committed as tests/golden/csv_sample.csv),
derived WITHOUT the restated front end -- a line-by-line Python restatement of the reference's Ragel
equivalent bench/ragel/src/csv2json.rl:9-31 (row = numVal ',' stringVal x4 ',' ipVal; the entering
and leaving actions print the key text, the quotes and the separators)."""
import os, re, sys
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "..", "tests", "golden")
KEYS = ["id", "first_name", "last_name", "email", "country", "ip"]
ROW = re.compile(rb"([0-9]+),([^,\n]*),([^,\n]*),([^,\n]*),([^,\n]*),([0-9]{1,3}\.[0-9]{1,3}\.[0-9]{1,3}\.[0-9]{1,3})\n")


def csv2json_rl(data: bytes) -> bytes:
    out = bytearray()
    for line in data.splitlines(keepends=True):
        m = ROW.fullmatch(line)
        assert m, line                                   # csv2json.rl: FAIL when p != pe
        out += b"{\n"                                    # csv2json = row > { P("{\n") }
        for i, key in enumerate(KEYS):
            val = m.group(i + 1)
            out += b"   " + b'"%s"' % key.encode() + b" " * (11 - len(key)) + b": "   # INDENT; P("\"id\"         : ")
            out += val if i == 0 else b'"' + val + b'"'  # numVal is echoed bare, stringVal / ipVal in quotes
            out += b",\n" if i < 5 else b"\n"            # % { OFF; P(",\n") }, the last field P("\n")
        out += b"}\n"                                    # % { P("}\n") }
    return bytes(out)


if __name__ == "__main__":
    data = open(os.path.join(GOLD, "csv_sample.csv"), "rb").read()
    out = csv2json_rl(data)
    open(os.path.join(GOLD, "csv2json_sample.out"), "wb").write(out)
    print(len(data), "->", len(out))
