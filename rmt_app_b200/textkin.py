"""Dashboard text format -> `reaction-rates` dict.

The PyREMOT dashboard lets users type kinetics as lines of  `"key": expression;`
(README.md:83-173 of the reference) and converts each line to
`"key": lambda x: expression`.  This module does that conversion; the resulting
dict is exactly what `rmtExe` expects and feeds the same tracer as live lambdas.

    VARS  = '''"CaBeDe": 1171.2; "RT": x['R_CONST']*x['T']; "K1": 35.45*math.exp(-1.7069e4/x['RT'])'''
    RATES = '''"r1": 1000*x['K1']*x['CaBeDe']'''
    reaction_rates = parse_reaction_rates(VARS, RATES)
"""
import ast
import math

import numpy as np

__all__ = ["parse_section", "parse_reaction_rates"]


def _split_entries(text):
    """Split on ';' or newlines that are not inside brackets/quotes."""
    out, depth, quote, cur = [], 0, None, []
    for ch in text:
        if quote:
            cur.append(ch)
            if ch == quote:
                quote = None
            continue
        if ch in "\"'":
            quote = ch
            cur.append(ch)
        elif ch in "([{":
            depth += 1
            cur.append(ch)
        elif ch in ")]}":
            depth -= 1
            cur.append(ch)
        elif ch == ";" and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return [e.strip() for e in out if e.strip()]


def parse_section(text, namespace=None):
    """`"key": expr; ...` -> ordered dict.  A numeric literal (or a name found in `namespace`
    that is numeric) becomes a scalar entry — a kinetic-parameter slot — everything else a
    `lambda x: expr` evaluated with `math` and `np` in scope."""
    env = {"math": math, "np": np, "numpy": np}
    env.update(namespace or {})
    out = {}
    for entry in _split_entries(text):
        key, sep, expr = entry.partition(":")
        if not sep:
            raise ValueError("entry %r is not of the form \"key\": expression" % entry)
        key = key.strip().strip("\"'")
        expr = " ".join(expr.split())
        if not key or not expr:
            raise ValueError("empty key or expression in %r" % entry)
        tree = ast.parse(expr, mode="eval")
        names = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name)}
        if "x" not in names:
            val = eval(compile(tree, "<kinetics:%s>" % key, "eval"), dict(env))
            if isinstance(val, (int, float, np.integer, np.floating)) and not isinstance(val, bool):
                out[key] = float(val)
                continue
        out[key] = eval(compile(ast.parse("lambda x: " + expr, mode="eval"), "<kinetics:%s>" % key, "eval"), dict(env))
    return out


def parse_reaction_rates(vars_text, rates_text, namespace=None):
    return {"VARS": parse_section(vars_text, namespace), "RATES": parse_section(rates_text, namespace)}
