// main.cpp — the reference's driver (src/main.rs:33-72) on the B200 engine:
//   tfhe_aes_cli --number-of-outputs N --iv X --key K
// key generation and encryption (client), AES key expansion, batched AES-128-CTR, decryption + verification.
#include <chrono>
#include <cstdio>
#include <cstring>
#include "tfhe_aes.hpp"

using namespace tfhe_aes;

static u128 parse_u128(const char *s) {
    u128 v = 0;
    if (s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) { for (s += 2; *s; s++) v = v * 16 + (u128)(*s <= '9' ? *s - '0' : (*s | 32) - 'a' + 10); return v; }
    for (; *s; s++) { if (*s < '0' || *s > '9') throw Error("bad number"); v = v * 10 + (u128)(*s - '0'); }
    return v;
}

int main(int argc, char **argv) {
    size_t number_of_outputs = 0;
    u128 iv = 0, key = 0;
    bool have_n = false;
    try {
        for (int i = 1; i + 1 < argc; i += 2) {  // main.rs:20-30
            if (!strcmp(argv[i], "--number-of-outputs")) { number_of_outputs = (size_t)parse_u128(argv[i + 1]); have_n = true; }
            else if (!strcmp(argv[i], "--iv")) iv = parse_u128(argv[i + 1]);
            else if (!strcmp(argv[i], "--key")) key = parse_u128(argv[i + 1]);
            else throw Error(std::string("unknown argument ") + argv[i]);
        }
        if (!have_n) throw Error("usage: tfhe_aes_cli --number-of-outputs N --iv X --key K");
        Engine engine(param_opt());
        Client client_obj(engine, number_of_outputs, iv, key);                       // main.rs:41
        auto [encrypted_iv, encrypted_key] = client_obj.client_encrypt();           // main.rs:43
        Server server_obj(engine);                                                   // main.rs:45
        auto t0 = std::chrono::steady_clock::now();
        RoundKeys encrypted_round_keys = server_obj.aes_key_expansion(encrypted_key);  // main.rs:49
        auto t1 = std::chrono::steady_clock::now();
        printf("AES key expansion took: %.3fs\n", std::chrono::duration<double>(t1 - t0).count());
        std::vector<u64> states = server_obj.aes_ctr(encrypted_round_keys, encrypted_iv, (int)number_of_outputs);  // main.rs:55-64
        auto t2 = std::chrono::steady_clock::now();
        printf("AES of #%zu outputs computed in: %.3fs\n", number_of_outputs, std::chrono::duration<double>(t2 - t1).count());
        client_decrypt_and_verify(client_obj, states);                               // main.rs:70
        printf("Passed: every output equals AES-128(key, iv + i).\n");
    } catch (const std::exception &ex) {
        fprintf(stderr, "error: %s\n", ex.what());
        return 1;
    }
    return 0;
}
