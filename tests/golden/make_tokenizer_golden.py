"""Builds tests/golden/llama3_style_tokenizer.json + tokenizer_golden.json with the HuggingFace
`tokenizers` library: a byte-level BPE trained on a small mixed corpus with EXACTLY the Llama-3
tokenizer pipeline (Split on the Llama-3 regex, ByteLevel without its own regex, ignore_merges,
the Llama-3 chat special tokens) — no Llama-3 vocabulary file exists in this image, so the PIPELINE
is pinned against the library that defines it, on a vocabulary small enough to commit.

    python tests/golden/make_tokenizer_golden.py
"""
import json
import os

from tokenizers import Regex, Tokenizer, decoders, models, pre_tokenizers, trainers

HERE = os.path.dirname(os.path.abspath(__file__))
LLAMA3_SPLIT = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*"
                r"|\s*[\r\n]+|\s+(?!\S)|\s+")
SPECIALS = ["<|begin_of_text|>", "<|end_of_text|>", "<|reserved_special_token_0|>", "<|reserved_special_token_1|>",
            "<|finetune_right_pad_id|>", "<|reserved_special_token_2|>", "<|start_header_id|>", "<|end_header_id|>",
            "<|eom_id|>", "<|eot_id|>", "<|python_tag|>"]

CORPUS = [
    "You are a helpful assistant. What is the capital of France? The capital of France is Paris.",
    "The moon is a natural satellite of the Earth and lacks any formal government or capital.",
    "I'm sure they've said we'll see; it's what he'd do, isn't it? DON'T SHOUT, I'M HERE. They'RE odd.",
    'Environment: ipython\n\nGiven the following functions, please respond with a JSON for a function call.\n\n'
    '{"type": "function", "function": {"name": "fetch__fetch", "description": "Fetch a URL", "parameters": '
    '{"type": "object", "properties": {"url": {"type": "string"}}, "required": ["url"]}}}\n\n',
    '{"name": "fetch__fetch", "parameters": {"url": "https://api.example.com/data?id=12345&q=abc"}}',
    "def reconcile(ctx, req):\n    task = get_task(req.namespace, req.name)\n    if task.status.phase == 'ReadyForLLM':\n"
    "        return send_llm_request(ctx, task)   # 1 HTTPS round trip\n\n\n    return Result(requeue_after=5)\n",
    "for (int i = 0; i < 1024; ++i) { sum += a[i] * b[i]; }\t// 3.14159 2718281828 0x7fff 1e-5 100000 12 1234567",
    "kubectl apply -f config/samples/acp_v1alpha1_task.yaml && kubectl get tasks -o wide --watch",
    "Ça va très bien, merci ! Ich heiße Jürgen und wohne in Köln. El niño comió piñas. Привет, мир! Γειά σου Κόσμε",
    "日本語のテキストと中文文本。한국어 텍스트도 있습니다。 العربية  עברית  हिन्दी ๑๒๓ ⅓ Ⅷ ½ ①②③",
    "emoji 🙂🚀👩‍💻 and symbols ±×÷≠≤≥ → ← ↑ ↓ … — – ‘quoted’ “double” «guillemets» ©®™ §¶ †‡ ° µ",
    "   leading spaces,  double  spaces,   triple   spaces   \n\n\nblank lines\r\nwindows line\r\n\r\n tab\there \n x",
    "x\u00a0y non-breaking\u2003em space\u2028line sep\u0085next line \u3000ideographic",
    "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA bbbbbbbbbbbbbbbbbbbbbbbb",
    "!!!???...,,,;;;:::---___***///\\\\\\|||~~~```^^^%%%$$$###@@@&&&+++===<<<>>>((()))[[[]]]{{{}}}",
] * 4

TESTS = CORPUS[:15] + [
    "", " ", "  ", "\n", " \n", "\n ", "a", " a", "a ", "  a", "a  b", "a   b", "a\nb", "a \nb", "a\n b", "a \n b", " \n\n ",
    "'s", "'S", "it's", "IT'S", "'re'll", "o'clock", "'", "''", "'x", "don't", "l'été", "123", "1234", "12345678", "1 2  3",
    "3.14", "v1.2.3-rc4", "abc123def", "123abc", "_under_score_", "camelCaseWord", "snake_case_word", "x=y+z;", "a.b.c()",
    "https://example.com/a/b?c=d&e=f#g", "user@example.com", "C:\\Users\\me\\file.txt", "/usr/local/bin/python3",
    "tab\tseparated\tvalues", "trailing space ", "trailing newline\n", "trailing  spaces  ", "\r\n", "a\r\nb\r\n\r\nc",
    "é", "éé", " é", "naïve café", "ÀÉÎÕÜ", "ß", "ǅ", "ʰ", "x²", "①", "१२३", "٣", "½", "Ⅻ", "a١b", "𝟘𝟙𝟚", "𝒜𝒷𝒸",
    "汉字", " 汉字", "汉字123", "かなカナ", "🙂", " 🙂", "🙂🙂", "a🙂b", "👩‍💻", "\u200b", "a\u200bb", "\ufeffbom", "\u00a0", " \u00a0 ",
    "x\u2028y", "x\u0085y", "\x0b\x0c", "\x00", "a\x00b", "\x7f", "\x1b[0m",
    "<|begin_of_text|>", "<|eot_id|> is just text here", "<|start_header_id|>user<|end_header_id|>",
    "The quick brown fox jumps over the lazy dog. " * 8,
    "word " * 50, "a" * 300, " " * 64, "\n" * 40, "ab" * 200, "9" * 50, "!" * 70,
]


def build():
    tok = Tokenizer(models.BPE(ignore_merges=True))
    tok.pre_tokenizer = pre_tokenizers.Sequence([
        pre_tokenizers.Split(Regex(LLAMA3_SPLIT), behavior="isolated", invert=False),
        pre_tokenizers.ByteLevel(add_prefix_space=False, trim_offsets=True, use_regex=False)])
    tok.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=1280, min_frequency=2, special_tokens=[],
                                  initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    tok.train_from_iterator(CORPUS, trainer)
    from tokenizers import AddedToken
    tok.add_special_tokens([AddedToken(s, special=True, normalized=False) for s in SPECIALS])
    return tok


def main():
    tok = build()
    path = os.path.join(HERE, "llama3_style_tokenizer.json")
    tok.save(path, pretty=False)
    tok = Tokenizer.from_file(path)
    cases = []
    for t in TESTS:
        # what the engine's encode() contract is: ordinary text, special spellings NOT recognised.
        # The library has no such switch on encode(), so specials are neutralised by encoding the
        # pieces around them — none of the TESTS except the three marked ones contain them.
        if "<|" in t:
            continue
        enc = tok.encode(t, add_special_tokens=False)
        cases.append({"text": t, "ids": enc.ids, "decoded": tok.decode(enc.ids, skip_special_tokens=False)})
    special = {s: tok.token_to_id(s) for s in SPECIALS}
    with open(os.path.join(HERE, "tokenizer_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_tokenizer_golden.py", "tokenizers_version": __import__("tokenizers").__version__,
                   "vocab_size": tok.get_vocab_size(with_added_tokens=True), "special": special, "cases": cases},
                  f, ensure_ascii=True)
    print(path, os.path.getsize(path), "bytes;", len(cases), "cases; vocab", tok.get_vocab_size(with_added_tokens=True))


if __name__ == "__main__":
    main()
