// C++ side of the BPE tokenizer (tokenizer.cc) for the serving shell: incremental detokenisation of a token stream.
#pragma once
#include <stdint.h>

#include <string>

struct b200_tokenizer;

namespace b200 {

// raw bytes one token id stands for (byte-level symbols mapped back; added tokens: their content, nothing for a special one
// when skip_special)
void tokenizer_token_bytes(const b200_tokenizer* t, int32_t id, bool skip_special, std::string* out);
// ids of the Llama-3 control tokens (-1 if the vocabulary lacks one)
int32_t tokenizer_bos(const b200_tokenizer* t);
int32_t tokenizer_eot(const b200_tokenizer* t);

// A token may end in the middle of a UTF-8 sequence (byte-level BPE): push() returns only text that is complete, holding an
// unfinished tail until the next token completes it; bytes that can never become valid are replaced by U+FFFD, as a lossy
// decode of the whole sequence would.  Concatenating every push() and the final flush() equals decoding all ids at once.
struct DetokStream {
  const b200_tokenizer* tok = nullptr;
  bool skip_special = true;
  std::string pending;
  std::string push(int32_t id);
  std::string flush();
};

}  // namespace b200
