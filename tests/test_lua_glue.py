"""Static checks of the LuaJIT drop-in (lua/bot7_b200/*.lua).  No Lua runtime exists in the image, so the glue cannot be
executed here; what CAN be checked on the CPU is that it is complete with respect to the reference's own call sites:

* every method the reference calls on the model / score / grid / sampler objects of the accelerated path
  (extracted from /root/reference when it is present, otherwise from the committed list below, which was extracted
  from it) is defined by the replacement class, or inherited from the reference parent it subclasses;
* every C symbol the glue calls through `B.C.` is declared in include/bot7_b200.h (and therefore in the generated cdef);
* block keywords balance (function/if/for/while/do ... end), the cheapest syntax check available without a parser.
"""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LUA = os.path.join(ROOT, "lua", "bot7_b200")
REF = "/root/reference"

# method names the reference calls on `self.model` / `model` in the files of the path:
#   bots/abstract.lua:148 init; bots/bayesopt.lua:65,68,74,75 class, sample_hypers, parse_hypers;
#   scores/expected_improvement.lua:57,63 and scores/confidence_bound.lua:57,63 fantasize, predict; models/abstract.lua cache
MODEL_METHODS = {"init", "sample_hypers", "parse_hypers", "class", "predict", "fantasize", "cache"}
DNGO_METHODS = {"init", "predict", "class"}            # bots/bayesopt.lua:65 + models/dngo.lua protocol used by the scores
CALL_SITES = ["bots/abstract.lua", "bots/bayesopt.lua", "scores/expected_improvement.lua", "scores/confidence_bound.lua"]


def lua(name):
    return open(os.path.join(LUA, name)).read()


def strip_comments(text):
    text = re.sub(r"--\[\[.*?\]\]", "", text, flags=re.S)
    return re.sub(r"--[^\n]*", "", text)


def defined_methods(text, cls):
    return set(re.findall(r"function\s+%s[:.]([A-Za-z_]\w*)\s*\(" % re.escape(cls), text))


def test_reference_call_sites_match_the_committed_list():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present (GPU box)")
    called = set()
    for f in CALL_SITES:
        src = strip_comments(open(os.path.join(REF, f)).read())
        called |= set(re.findall(r"\bmodel:([A-Za-z_]\w*)\s*\(", src))
    assert called <= MODEL_METHODS, f"the reference calls model methods missing from the list: {called - MODEL_METHODS}"
    assert {"init", "sample_hypers", "parse_hypers", "class", "predict", "fantasize"} <= called


def test_gp_regressor_defines_every_method_the_reference_calls():
    got = defined_methods(lua("models_gp.lua"), "model")
    assert MODEL_METHODS <= got, f"missing: {MODEL_METHODS - got}"
    # the density the sampler sees carries the prior and the failure guard, like the Python twin
    text = lua("models_gp.lua")
    assert "prior_std" in text and "-math.huge" in text and "info[" in text


def test_dngo_override_defines_the_hand_off():
    text = lua("models_dngo.lua")
    got = defined_methods(text, "dngo")
    assert {"predict", "class", "acquire", "basis_stack"} <= got
    assert "'bot7.models.dngo'" in text                      # subclasses the reference class: init / report / network stay the parent's
    for sym in ("b7_mlp_features", "b7_blr_fit", "b7_blr_predict", "b7_dngo_score"):
        assert "B.C.%s" % sym in text


def test_bayesopt_calls_only_defined_model_methods_and_uses_nGPU():
    text = strip_comments(lua("bayesopt.lua"))
    called = set(re.findall(r"self\.model:([A-Za-z_]\w*)\s*\(", text))
    gp = defined_methods(lua("models_gp.lua"), "model")
    dn = defined_methods(lua("models_dngo.lua"), "dngo")
    for m in called:
        assert m in gp or m in dn, f"bayesopt.lua calls model:{m}, which no replacement model defines"
    assert "config.bot.nGPU" in text
    for sym in ("b7_comm_init_all", "b7_gp_fit_sharded", "b7_acq_score_multi", "b7_grid_remove_sharded"):
        assert "B.C.%s" % sym in text


def test_glue_calls_only_declared_symbols():
    header = open(os.path.join(ROOT, "include", "bot7_b200.h")).read()
    declared = set(re.findall(r"\b(b7_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S)))
    enums = set(re.findall(r"\b(B7_[A-Z0-9_]+)\b", header))
    for f in os.listdir(LUA):
        if not f.endswith(".lua") or f == "ffi.lua":
            continue
        text = strip_comments(lua(f))
        for sym in set(re.findall(r"B\.C\.(b7_[a-z0-9_]+)", text)):
            assert sym in declared, f"{f} calls {sym}, not declared in include/bot7_b200.h"
        for e in set(re.findall(r"B\.C\.(B7_[A-Z0-9_]+)", text)):
            assert e in enums, f"{f} uses {e}, not an enum of include/bot7_b200.h"


def test_install_replaces_every_accelerated_class():
    text = lua("init.lua")
    for target in ("bot7.grids.sobol", "bot7.scores.expected_improvement", "bot7.scores.confidence_bound", "bot7.models.gp_regressor",
                   "bot7.models.dngo", "bot7.bots.bayesopt"):
        assert re.search(re.escape(target) + r"\s*=", text), f"install() does not replace {target}"


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(LUA) if f.endswith(".lua")))
def test_block_keywords_balance(name):
    text = strip_comments(lua(name))
    text = re.sub(r"\[\[.*?\]\]", "", text, flags=re.S)           # long strings (the cdef)
    text = re.sub(r"'(?:[^'\\\n]|\\.)*'|\"(?:[^\"\\\n]|\\.)*\"", "''", text)
    opens = len(re.findall(r"\bfunction\b", text)) + len(re.findall(r"\bif\b", text)) + len(re.findall(r"\bdo\b", text))
    # `for ... do` and `while ... do` are counted once through their `do`; `repeat ... until` has no `end`
    ends = len(re.findall(r"\bend\b", text))
    assert opens == ends, f"{name}: {opens} block openers vs {ends} `end`"
    assert text.count("(") == text.count(")") and text.count("{") == text.count("}")
