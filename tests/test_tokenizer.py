"""The byte-level BPE tokenizer (kubeai_b200/csrc/tokenizer.cc) against the HF `tokenizers` wheel — the library the
reference's backend uses on the checkpoint's tokenizer.json — on a tokenizer.json of the Llama-3 pipeline built here
(no network: the real Llama-3 vocabulary cannot be fetched; the pipeline, the pre-tokenizer pattern, ignore_merges, the
special tokens and the chat template are Llama-3's, the merges are trained on a small multilingual corpus).  Exact id
equality; CPU only."""
import json
import random

import pytest

tokenizers = pytest.importorskip("tokenizers")
from tokenizers import Regex, Tokenizer as HFTokenizer, decoders, models, pre_tokenizers, trainers  # noqa: E402

PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")
SPECIALS = ["<|begin_of_text|>", "<|end_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>", "<|reserved_special_token_0|>"]
CORPUS = [
    "The quick brown fox jumps over the lazy dog. It's the tokenizer's job; we'll see if they're right, I'd say I'm sure you've",
    "def encode(self, text: str) -> list[int]:\n    return [self.vocab[t] for t in text.split()]  # 12345 67 890\n\n\n",
    "naïve café déjà vu, Ünïcödé straße — ¿qué tal? Привет, мир! Γειά σου "
    "Κόσμε. שלום עולם مرحبا بالعالم",
    "你好，世界！今天天气很好。日本語のテキストも含まれています。"
    "한국어 문장도 있습니다. \U0001f642\U0001f680\U0001f44d\U0001f3fd ①②③ Ⅷ ½ ٣٤٥",
    "   leading spaces\tand\ttabs \n trailing   \r\n mixed\r\rnewlines \x0b\x0c  end  ",
    "https://example.com/path?query=1&x=%20y user@example.org {\"key\": [1, 2.5, -3e10], \"s\": 'q'} <html>&amp;</html>",
    "I'M SHOUTING AND IT'S 'RE 'VE 'LL 'D 'S 'T 'M in CAPS, don't can't won't they'll we'd",
] * 30


@pytest.fixture(scope="module")
def pair(tmp_path_factory):
    tok = HFTokenizer(models.BPE(ignore_merges=True))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(PATTERN), behavior="isolated", invert=False),
                                                 pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    tok.decoder = decoders.ByteLevel()
    tr = trainers.BpeTrainer(vocab_size=3000, special_tokens=SPECIALS, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    tok.train_from_iterator(CORPUS, tr)
    path = tmp_path_factory.mktemp("tok") / "tokenizer.json"
    tok.save(str(path))
    from kubeai_b200.tokenizer import Tokenizer
    mine = Tokenizer(path)
    yield tok, mine, path
    mine.close()


CASES = CORPUS[:7] + [
    "", " ", "  ", "\n", " \n", "\n ", "a", " a", "  a", "a ", "a  ", "a\n\nb", "a \n b", "x  \n  y", "\t\tx", " \t x", "\r\n", "a\r\nb",
    "123", "1234", "1234567", "12 345", "1,234.56", "a1b2c3", "٣٤٥٦٧", "x²", "①②③④",
    "'s", "'S", "it's", "IT'S", "'re'", "'", "''", "'x", "'lll", "'ve been", "d'accord", "ſ'ſ", "it'ſ", "K'K",
    "!!!", " !!!", "  !!!", "!!!\n\n", "?!\r\n\r\nx", "a-b", "a - b", "--flag=value", "( )", "()", " ()\n",
    "<|eot_id|>", "x<|eot_id|>y", "<|eot_id|><|eot_id|>", "<|eot_id", "<|start_header_id|>user<|end_header_id|>\n\nhi<|eot_id|>",
    "\U0001f642", " \U0001f642", "\U0001f642\U0001f642 x", "\U0001f44d\U0001f3fd", "á", " x", "x　y", " ", "", "e​f", "﻿bom",
    "ÀÉÎõü", "ßẞ", "Ǆǅǆ", "\U0001d4b3\U0001d4b4", "한국어", "日本語テキスト",
    "عربى ١٢٣", "हिन्दी १२३",
]


def test_ids_equal_hf_tokenizers_on_the_case_list(pair):
    hf, mine, _ = pair
    for s in CASES:
        want = hf.encode(s).ids
        got = mine.encode(s)
        assert got == want, (s, got, want, hf.encode(s).tokens)
        assert mine.decode(got) == hf.decode(want, skip_special_tokens=False), s


def test_ids_equal_hf_tokenizers_on_random_text(pair):
    hf, mine, _ = pair
    rng = random.Random(7)
    pools = [
        [chr(c) for c in range(0x20, 0x7F)] + list("\n\r\t "),
        [chr(c) for c in range(0xA0, 0x250)],
        [chr(c) for c in list(range(0x370, 0x400)) + list(range(0x400, 0x460)) + list(range(0x5D0, 0x5EB)) + list(range(0x620, 0x66A))],
        [chr(c) for c in list(range(0x3040, 0x30A0)) + list(range(0x4E00, 0x4E80)) + list(range(0xAC00, 0xAC40))],
        [chr(c) for c in list(range(0x2000, 0x2070)) + list(range(0x2150, 0x2190)) + list(range(0x2460, 0x2480)) + list(range(0x1F600, 0x1F640))],
        list(" \n\t'0123456789") + ["'s", "'RE", "  ", "\r\n", "<|eot_id|>"],
    ]
    for i in range(600):
        k = rng.randint(1, 40)
        s = "".join(rng.choice(rng.choice(pools)) for _ in range(k))
        want = hf.encode(s).ids
        got = mine.encode(s)
        assert got == want, (i, repr(s), got, want)
    # code points next to a letter, a digit and a space: the class tables and the alternation order
    for cp in list(range(0, 0x3000, 1)) + list(range(0x3000, 0x10000, 7)) + list(range(0x10000, 0x20000, 97)):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        s = f"a{chr(cp)}b 1{chr(cp)}2 {chr(cp)}  {chr(cp)}\n"
        assert mine.encode(s) == hf.encode(s).ids, hex(cp)


def test_special_tokens_are_parsed_or_left_as_text(pair):
    hf, mine, _ = pair
    s = "a<|eot_id|>b"
    assert mine.encode(s, allow_special=True) == hf.encode(s).ids
    plain = mine.encode(s, allow_special=False)
    assert mine.token_id("<|eot_id|>") not in plain and mine.decode(plain) == s
    ids = mine.encode("x<|begin_of_text|>y")
    assert mine.decode(ids, skip_special=True) == "xy" and mine.decode(ids) == "x<|begin_of_text|>y"
    assert mine.vocab_size == hf.get_vocab_size(with_added_tokens=True)


LLAMA3_TEMPLATE = ("{% set loop_messages = messages %}{% for message in loop_messages %}{% set content = '<|start_header_id|>' + message['role'] + "
                   "'<|end_header_id|>\n\n'+ message['content'] | trim + '<|eot_id|>' %}{% if loop.index0 == 0 %}{% set content = bos_token + content %}"
                   "{% endif %}{{ content }}{% endfor %}{% if add_generation_prompt %}{{ '<|start_header_id|>assistant<|end_header_id|>\n\n' }}{% endif %}")


def test_chat_template_equals_transformers_apply_chat_template(pair):
    transformers = pytest.importorskip("transformers")
    hf, mine, _ = pair
    fast = transformers.PreTrainedTokenizerFast(tokenizer_object=hf, bos_token="<|begin_of_text|>", eos_token="<|eot_id|>", chat_template=LLAMA3_TEMPLATE)
    convs = [
        [{"role": "user", "content": "Hello there!"}],
        [{"role": "system", "content": "  You are a helpful assistant.\n"}, {"role": "user", "content": "What's 12345 + 678?\n\n"},
         {"role": "assistant", "content": "It's 13023."}, {"role": "user", "content": "谢谢 \U0001f642  "}],
        [{"role": "user", "content": ""}],
    ]
    for msgs in convs:
        for gen in (True, False):
            want = fast.apply_chat_template(msgs, tokenize=True, add_generation_prompt=gen)
            want = want["input_ids"] if hasattr(want, "keys") else want
            assert mine.chat(msgs, add_generation_prompt=gen) == list(want), (msgs, gen)


def test_other_pipelines_are_refused_not_approximated(pair, tmp_path):
    from kubeai_b200 import B200Error
    from kubeai_b200.tokenizer import Tokenizer
    _, _, path = pair
    bad = json.loads(path.read_text())
    bad["pre_tokenizer"] = {"type": "ByteLevel", "add_prefix_space": False, "trim_offsets": True, "use_regex": True}   # GPT-2's own regex
    (tmp_path / "gpt2.json").write_text(json.dumps(bad))
    with pytest.raises(B200Error, match="unsupported tokenizer"):
        Tokenizer(tmp_path / "gpt2.json")
    bad = json.loads(path.read_text())
    bad["normalizer"] = {"type": "NFC"}
    (tmp_path / "nfc.json").write_text(json.dumps(bad))
    with pytest.raises(B200Error, match="normalizer"):
        Tokenizer(tmp_path / "nfc.json")
    with pytest.raises(B200Error):
        Tokenizer(tmp_path / "missing.json")


def test_incremental_detokenisation_never_emits_a_broken_utf8_sequence(pair):
    from kubeai_b200.tokenizer import DetokStream
    hf, mine, _ = pair
    rng = random.Random(3)
    texts = ["naïve café 你好 \U0001f642\U0001f44d\U0001f3fd done", "日本語のテキスト ①②③ ½", "plain ascii only", "\U0001f680" * 5]
    seqs = [mine.encode(t) for t in texts]
    seqs += [[rng.randrange(6, mine.vocab_size) for _ in range(60)] for _ in range(40)]     # random ids: arbitrary byte soup
    for ids in seqs:
        st = DetokStream(mine, skip_special=False)
        out = "".join(st.push(i) for i in ids) + st.flush()        # push() decodes strictly: raises on a broken sequence
        st.close()
        assert out == mine.decode(ids)
        assert out == hf.decode(ids, skip_special_tokens=False)
    st = DetokStream(mine, skip_special=True)
    ids = mine.encode("a<|eot_id|>b")
    assert "".join(st.push(i) for i in ids) + st.flush() == "ab"
    st.close()


def test_server_renders_prompts_with_the_attached_tokenizer(pair):
    transformers = pytest.importorskip("transformers")
    from kubeai_b200.server import Server
    hf, mine, _ = pair
    fast = transformers.PreTrainedTokenizerFast(tokenizer_object=hf, bos_token="<|begin_of_text|>", eos_token="<|eot_id|>", chat_template=LLAMA3_TEMPLATE)
    with Server([None], model="m", vocab=4096) as srv:
        msgs = [{"role": "system", "content": "Be brief."}, {"role": "user", "content": "What's up? 你好 12345"}]
        body = json.dumps({"model": "m", "messages": msgs})
        synthetic = srv.render_prompt("/v1/chat/completions", body)
        srv.set_tokenizer(mine)
        want = fast.apply_chat_template(msgs, tokenize=True, add_generation_prompt=True)
        want = want["input_ids"] if hasattr(want, "keys") else want
        assert srv.render_prompt("/v1/chat/completions", body) == list(want) != synthetic
        # a conversation that already ends with an assistant turn is continued, not re-opened
        cont = msgs + [{"role": "assistant", "content": "Not much"}]
        want = fast.apply_chat_template(cont, tokenize=True, add_generation_prompt=False)
        want = want["input_ids"] if hasattr(want, "keys") else want
        assert srv.render_prompt("/v1/chat/completions", json.dumps({"model": "m", "messages": cont})) == list(want)
        # completions: <|begin_of_text|> + the text, as vLLM's add_special_tokens default does for this family
        text = "Once upon a time, in 1999"
        assert srv.render_prompt("/v1/completions", json.dumps({"model": "m", "prompt": text})) == [mine.token_id("<|begin_of_text|>")] + hf.encode(text).ids
        assert srv.render_prompt("/v1/completions", json.dumps({"model": "m", "prompt": [5, 6, 7]})) == [5, 6, 7]
        srv.set_tokenizer(None)
        assert srv.render_prompt("/v1/chat/completions", body) == synthetic
    from kubeai_b200 import B200Error
    with Server([None], model="m", vocab=300) as small:
        with pytest.raises(B200Error, match="vocabulary"):
            small.set_tokenizer(mine)
