"""serde_yaml 0.9 string scalars: the product's emitter (blu_decode.h yaml_str, compiled for the host by tests/sim_ffi)
against the independent Python statement of the same rules (oracle/pyoracle.py yaml_str), byte for byte, on hand-picked
and random strings.  serde_yaml itself is not vendored under /root/reference: these rules are a restatement (unpinned)."""
import random

import pyoracle as po
import sim_ffi

HAND = ["abc", "123", "007", "-007", "1e5", "0x1F", "0xZZ", "-5", "+5", "++5", "+-5", "yes", "Yes", "YES", "oFf", "~", "", "a: b", "a:", "a :b", "a #b", "a#b",
        "#x", "- x", "-x", "-", "- ", " x", "x ", "it's", "'q'", "tab\there", "1.", "1e999", "-1e999", ".5", "5.", ".", "e5", "1e", "1e+", "d__bacteria;p__x",
        "NR_1.1", "true", "tRuE", "Null", "nULL", "...x", "---", "--", "?", "? x", "?x", ":x", ": x", ":", "é", "ünï", "1_000", "--5", "0o17", "0o8",
        "0b2", "0b101", "inf", ".inf", "-.INF", ".NaN", "nan", "+.inf", "340282366920938463463374607431768211455", "340282366920938463463374607431768211456",
        "99999999999999999999999999999999999999999999", "0x" + "F" * 32, "0x" + "F" * 33, "[x]", "{x", "&a", "*a", "!a", "|a", ">a", "%a", "@a", "`a", "a,b", ",a",
        "a\x01b", "\x7f", "ab", "a b", "﻿x", "a b", "q\\x", 'q"x', "q'x y: z", "x\ty", "trailing:", "a  #  b", "12abc", "1.2.3", "1-2", "0",
        "00", "0.0", "-0", "+0", "0e0", "1E5", "1e-5", "1e+5", "SRR1.5_size_3", "NR_100000.1", "\U0001F600", "a\x1bb", "a\x00b"]


def test_hand_picked():
    for s in HAND:
        assert sim_ffi.yaml_str(s) == po.yaml_str(s), repr(s)


def test_random_strings():
    rng = random.Random(8)
    alphabet = "0123456789abxXoOeE.+-_:# '\"~ynYN?-[]{},&*!|>%@`\té\x01"
    for _ in range(20000):
        s = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 7)))
        assert sim_ffi.yaml_str(s) == po.yaml_str(s), repr(s)
