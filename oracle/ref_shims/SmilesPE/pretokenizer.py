"""SmilesPE.pretokenizer.atomwise_tokenizer restated (import at MolNexTR/tokenization.py:6;
only the training path calls it)."""
import re

_PATTERN = r"(\[[^\]]+]|Br?|Cl?|N|O|S|P|F|I|b|c|n|o|s|p|\(|\)|\.|=|#|-|\+|\\|\/|:|~|@|\?|>|\*|\$|\%[0-9]{2}|[0-9])"


def atomwise_tokenizer(smi, exclusive_tokens=None):
    return [t for t in re.compile(_PATTERN).findall(smi)]
