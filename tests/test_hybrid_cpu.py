"""`-hybrid` host graphs. The reference reads a GML hybridisation network, validates it and then fails on
its first guest vertex: GuestTreeInHybridGraphCreator.java:272 passes `null` for the guest tree and GuestVertex.java:100
dereferences it. A drop-in therefore ends every such run like the reference does - the message of the first check that fails, or
the NullPointerException for a valid network, then "Use option -h ...", no files. These tests compare the product's host-side
argument path (`sgp_probe_parse_app`, no GPU) with the oracle's restatement (oracle/hybrid.hpp) on every branch of the checks."""
import os

import pytest

import sagephy_b200 as sg
from oracle_util import oracle_run

TAIL = "\n\nUse option -h or --help to show usage.\n"

# stem tip 0, speciation 1, hybrid donors 2 and 3, allopolyploid hybrid 4, leaves A, B, H
ALLO = """# allopolyploid hybrid of two lineages
graph [
  directed 1
  name "net"
  node [ id 0 time 2.0 ]
  node [ id 1 time 1.5 ]
  node [ id 2 time 1.0 ]
  node [ id 3 time 1.0 ]
  node [ id 4 time 1.0 ]
  node [ id 5 name "A" time 0 ]
  node [ id 6 name "B" time 0.0 graphics [ x 1.0 y 2.0 ] ]
  node [ id 7 name "H" time 0.0 ]
  edge [ source 0 target 1 ]
  edge [ source 1 target 2 ]
  edge [ source 1 target 3 ]
  edge [ source 2 target 4 ]
  edge [ source 3 target 4 ]
  edge [ source 2 target 5 ]
  edge [ source 3 target 6 ]
  edge [ source 4 target 7 ]
]
"""
# stem tip 0, hybrid donor 1 (leaf X below it), autopolyploid hybrid 2 on a zero-length arc from 1, leaf Y below the hybrid
AUTO = """graph [ directed 1
  node [ id 0 time 3.0 ] node [ id 1 time 2.0 ] node [ id 2 time 2.0 ] node [ id 3 name "X" time 0.0 ] node [ id 4 name "Y" time 0.0 ]
  edge [ source 0 target 1 ] edge [ source 1 target 2 ] edge [ source 1 target 3 ] edge [ source 2 target 4 ]
]"""


def swap(text, old, new):
    assert old in text
    return text.replace(old, new, 1)


BAD_GRAPHS = {
    "not_directed": swap(ALLO, "directed 1", "directed 0"),
    "directed_missing": swap(ALLO, "  directed 1\n", ""),
    "no_graph": "Creator \"x\" Version 1",
    "vertex_lacks_id": swap(ALLO, "node [ id 2 time 1.0 ]", "node [ time 1.0 ]"),
    "duplicate_id": swap(ALLO, "node [ id 2 time 1.0 ]", "node [ id 1 time 1.0 ]"),
    "id_out_of_range": swap(ALLO, "node [ id 2 time 1.0 ]", "node [ id 12 time 1.0 ]"),
    "no_time": swap(ALLO, "node [ id 2 time 1.0 ]", "node [ id 2 ]"),
    "time_not_a_number": swap(ALLO, "node [ id 2 time 1.0 ]", "node [ id 2 time \"1.0\" ]"),
    "arc_target_out_of_range": swap(ALLO, "edge [ source 4 target 7 ]", "edge [ source 4 target 9 ]"),
    "arc_backwards_in_time": swap(ALLO, "edge [ source 2 target 5 ]", "edge [ source 5 target 2 ]"),
    "self_loop": ALLO.replace("]\n]", "]\n  edge [ source 1 target 1 ]\n]"),
    "edge_without_target": swap(ALLO, "edge [ source 4 target 7 ]", "edge [ id 3 source 4 ]"),
    "not_connected": ALLO.replace("edge [ source 4 target 7 ]\n", "").replace("node [ id 7 name \"H\" time 0.0 ]", "node [ id 7 name \"H\" time 0.0 ]\n  node [ id 8 name \"Z\" time 0.0 ]\n  edge [ source 7 target 8 ]"),
    "stem_tip_with_two_children": ALLO.replace("]\n]", "]\n  edge [ source 0 target 2 ]\n]"),
    "leaf_on_zero_arc": swap(ALLO, "node [ id 7 name \"H\" time 0.0 ]", "node [ id 7 name \"H\" time 1.0 ]"),
    "unary_vertex_between_long_arcs": swap(AUTO, "node [ id 2 time 2.0 ]", "node [ id 2 time 1.0 ]"),
    "two_zero_children": swap(swap(ALLO, "node [ id 5 name \"A\" time 0 ]", "node [ id 5 name \"A\" time 1.0 ]"), "edge [ source 2 target 5 ]", "edge [ source 2 target 5 ]"),
    "hybrid_with_a_long_parent_arc": swap(ALLO, "node [ id 3 time 1.0 ]", "node [ id 3 time 1.2 ]"),
    "three_parents": ALLO.replace("]\n]", "]\n  edge [ source 1 target 4 ]\n]"),
    "leaf_without_name": swap(ALLO, "node [ id 7 name \"H\" time 0.0 ]", "node [ id 7 time 0.0 ]"),
    "leaf_with_empty_name": swap(ALLO, "name \"H\"", "name \"\""),
    "syntax_missing_bracket": ALLO.rstrip()[:-1],
    "syntax_trailing": ALLO + "]",
    "syntax_bad_key": swap(ALLO, "name \"net\"", "9name \"net\""),
    "syntax_bad_number": swap(ALLO, "time 1.5", "time 1.5.2"),
    "syntax_open_string": swap(ALLO, "name \"H\"", "name \"H"),
    "type_of_id": swap(ALLO, "node [ id 2 time 1.0 ]", "node [ id \"2\" time 1.0 ]"),
    "type_of_source": swap(ALLO, "edge [ source 4 target 7 ]", "edge [ source 4.0 target 7 ]"),
    "type_of_directed": swap(ALLO, "directed 1", "directed \"yes\""),
    # lines are queued without their line breaks: "time 1.5" + "edge" reads as the number 1.5e
    "token_runs_into_next_line": "graph [\ndirected 1\nnode [ id 0 time 1.5\nedge 1 ]\n]",
    "node_is_not_a_list": swap(ALLO, "node [ id 7 name \"H\" time 0.0 ]", "node \"H\""),
    "graph_is_not_a_list": "graph 5 " + ALLO,
    # GMLFactory.getGraphs builds every graph before the first one is used
    "second_graph_is_broken": ALLO + "graph [ node [ id \"x\" ] ]",
}


def both(app, args):
    rc, text = sg.probe_parse_app(app, args)
    want = oracle_run(app, args)
    assert want["files"] == {} and want["stdout"] == ""
    return rc, text, want["stderr"]


@pytest.mark.parametrize("text", [ALLO, AUTO], ids=["allopolyploid", "autopolyploid"])
@pytest.mark.parametrize("app", ["GuestTreeGen", "DomainTreeGen"])
def test_valid_hybrid_network_ends_like_the_reference(tmp_path, app, text):
    f = tmp_path / "net.gml"
    f.write_text(text)
    rates = ["0.3", "0.2", "0.0"]
    for host in (str(f), text):                                   # a file, or the GML text itself as the argument
        args = ["-s", "4", "-hybrid", "0.2", "2.0", "0.5", host] + (["genes"] if app == "DomainTreeGen" else []) + rates + [str(tmp_path / "out")]
        rc, got, want = both(app, args)
        assert rc != 0 and got == want == "java.lang.NullPointerException" + TAIL
    assert sorted(os.listdir(tmp_path)) == ["net.gml"]


@pytest.mark.parametrize("name", sorted(BAD_GRAPHS))
def test_hybrid_network_checks_match_the_oracle(name):
    args = ["-s", "4", "-hybrid", "0.2", "2.0", "0.5", BAD_GRAPHS[name], "0.3", "0.2", "0.0", "out"]
    rc, got, want = both("GuestTreeGen", args)
    assert rc != 0 and got == want, (name, got, want)
    assert got.endswith(TAIL) and not got.startswith("java.lang.NullPointerException" + TAIL) or name == "directed_missing"


@pytest.mark.parametrize("args,needle", [
    (["-s", "4", "-hybrid", "0.2", "2.0", "x", ALLO, "0.3", "0.2", "0.0", "out"], "NumberFormatException"),
    (["-s", "4", "-hybrid", "0.2", "2.0", "0.5", ALLO, "zz", "0.2", "0.0", "out"], "NumberFormatException"),
    (["-s", "4", "-hybrid", "-0.2", "2.0", "0.5", ALLO, "0.3", "0.2", "0.0", "out"], "Cannot have rate or rate factor less than 0."),
    (["-s", "4", "-hybrid", "0.2", "2.0", "0.5", ALLO, "-0.3", "0.2", "0.0", "out"], "Cannot have rate or rate factor less than 0."),
    (["-s", "4", "-hybrid", "0.2", "2.0", "0.5", "-p", "1.5", ALLO, "0.3", "0.2", "0.0", "out"], "Cannot have leaf sampling probability outside [0,1]."),
    (["-s", "4", "-hybrid", "0.2", "2.0", "0.5", "-p", "p", ALLO, "0.3", "0.2", "0.0", "out"], "NumberFormatException"),
    (["-s", "4", "-hybrid", "0.2", "2.0", "0.5", "-stem", "s", ALLO, "0.3", "0.2", "0.0", "out"], "NumberFormatException"),
    (["-s", "4", ALLO, "0.3", "0.2", "0.0", "out", "-hybrid", "0.2", "2.0"], "Expected a value after parameter -hybrid"),
])
def test_hybrid_argument_checks_match_the_oracle(args, needle):
    rc, got, want = both("GuestTreeGen", args)
    assert rc != 0 and needle in got and needle in want
    assert got.endswith(TAIL) and want.endswith(TAIL)


def test_parse_probe_accepts_ordinary_runs(tmp_path):
    host = "((A:0.4,B:0.4):0.6,C:1.0):0.2;"
    assert sg.probe_parse_app("GuestTreeGen", ["-s", "1", host, "0.3", "0.3", "0.1", "o"]) == (0, "")
    assert sg.probe_parse_app("HostTreeGen", ["-s", "1", "1.0", "2.0", "0.3", "o"]) == (0, "")
    rc, text = sg.probe_parse_app("HostTreeGen", ["1.0", "2.0"])
    assert rc == sg.STATUS_USAGE and text == ""
    rc, text = sg.probe_parse_app("GuestTreeGen", ["-s", "1", host, "-0.3", "0.3", "0.1", "o"])
    assert rc != 0 and "Cannot have rate less than 0." in text


def test_mutated_networks_get_the_same_answer_from_product_and_oracle():
    # two independent readers of the same grammar (a one-pass scanner in the product, a key-value tree like the Java in the oracle) on
    # random edits of valid networks: deleted spans, inserted and substituted tokens
    import random
    rng = random.Random(20261020)
    tokens = ['[', ']', '"', ' ', '\n', 'id', 'time', 'node', 'edge', 'source', 'target', 'directed', 'graph', 'name', '1', '0', '7', '1.5', '-', 'e', '.',
              'label', 'graphics', '#', '2.0', '9', 'x']
    checked = 0
    while checked < 1500:
        t = list(rng.choice([ALLO, AUTO]))
        for _ in range(rng.choice([1, 1, 2, 3])):
            k = rng.randrange(len(t)); op = rng.random()
            if op < 0.35:
                del t[k:k + rng.choice([1, 1, 2, 5, 12])]
            elif op < 0.7:
                t[k:k] = list(rng.choice(tokens))
            else:
                t[k:k + rng.choice([1, 3])] = list(rng.choice(tokens))
        text = "".join(t)
        if not text.strip():
            continue
        args = ["-s", "4", "-hybrid", "0.2", "2.0", "0.5", text, "0.3", "0.2", "0.0", "out"]
        rc, got = sg.probe_parse_app("GuestTreeGen", args)
        want = oracle_run("GuestTreeGen", args)
        assert rc != 0 and got == want["stderr"] and want["files"] == {}, (text, got, want["stderr"])
        checked += 1


def _gml_syntax_ok(text):
    """Third, independent reader of the grammar of io/GMLReader.java:62-231 (Python, syntax only): does the text parse?"""
    ws = set(" \t\n\x0b\x0c\r\x1c\x1d\x1e\x1f")
    lines = text.strip(''.join(chr(c) for c in range(33))).split("\n")
    kept = []
    for ln in lines:
        first = next((c for c in ln if c not in ws), None)
        if first != "#":
            kept.append(ln)
    q = "".join(kept); n = len(q); pos = 0

    def skip():
        nonlocal pos
        while pos < n and q[pos] in ws:
            pos += 1

    def plist():
        nonlocal pos
        while True:
            skip()
            if pos >= n or q[pos] == "]":
                return True
            if not (q[pos].isascii() and q[pos].isalpha()):
                return False
            pos += 1
            while pos < n and q[pos].isascii() and q[pos].isalnum():
                pos += 1
            skip()
            if pos < n and q[pos] == "[":
                pos += 1
                if not plist() or pos >= n or q[pos] != "]":
                    return False
                pos += 1
            elif pos < n and q[pos] == '"':
                end = q.find('"', pos + 1)
                if end < 0:
                    return False
                pos = end + 1
            else:
                start = pos
                while pos < n and q[pos] in "0123456789.Ee+-":
                    pos += 1
                tok = q[start:pos]
                try:
                    if "." in tok:
                        float(tok)
                        if tok.lower().lstrip("+-").startswith(("inf", "nan")):
                            return False
                    else:
                        if not tok or not tok.lstrip("+-").isdigit() or len(tok) - len(tok.lstrip("+-")) > 1 or not -2 ** 31 <= int(tok) < 2 ** 31:
                            return False
                except ValueError:
                    return False
    return plist() and pos >= n


def test_gml_syntax_verdict_agrees_with_an_independent_reader():
    import random
    rng = random.Random(7)
    tokens = ['[', ']', '"', ' ', '\n', 'id', 'time', 'node', 'edge', 'graph', '1', '1.5', '-', 'e', '.', '#', 'x', '+', '2e3', '.5']
    for k in range(1200):
        t = list(rng.choice([ALLO, AUTO]))
        for _ in range(rng.choice([1, 2, 3])):
            i = rng.randrange(len(t)); op = rng.random()
            if op < 0.4:
                del t[i:i + rng.choice([1, 2, 7])]
            else:
                t[i:i] = list(rng.choice(tokens))
        text = "".join(t)
        if not text.strip():
            continue
        err = oracle_run("GuestTreeGen", ["-s", "4", "-hybrid", "0.2", "2.0", "0.5", text, "0.3", "0.2", "0.0", "out"])["stderr"]
        if err.startswith("ParameterException"):          # the mutated text begins with '-' and is taken for an option
            continue
        syntax_error = err.startswith("uc.sgp.sagephy.io.GMLIOException: Error parsing GML file.")
        assert syntax_error == (not _gml_syntax_ok(text)), (text, err[:200])
