def sent_tokenize(*a, **k):
    raise NotImplementedError


def word_tokenize(*a, **k):
    raise NotImplementedError
