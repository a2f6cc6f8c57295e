// exp(x) for Float64 exactly as Julia computes it (base/special/exp.jl: 256-entry table of 2^(j/256) split into a rounded-down head and
// a 12-bit packed tail, degree-3 minimax kernel for expm1 on |r| <= ln2/512, fused multiply-adds throughout).  Neither Julia's nor
// glibc's exp is correctly rounded, and they disagree in the last bit for a fraction of a per cent of the arguments; the reference's
// grid-sampled initial conditions (bkw / maxwellian evaluated on a velocity lattice) feed those bits into the octree merge's choice
// between mirror-image bins, so both the CPU oracle and the device library's host-side weight table evaluate the distribution with
// this restatement.  The table is regenerated from its definition (tests/golden/make_jlexp_table.py); normal range only.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace mbjl {

static const uint64_t J_TABLE[256] = {
    0x0000000000000000ull, 0xaac00b1afa5abcbeull, 0x9b60163da9fb3335ull, 0xab502168143b0280ull,
    0xadc02c9a3e778060ull, 0x656037d42e11bbccull, 0xa7a04315e86e7f84ull, 0x84c04e5f72f654b1ull,
    0x8d7059b0d3158574ull, 0xa510650a0e3c1f88ull, 0xa8d0706b29ddf6ddull, 0x83207bd42b72a836ull,
    0x6180874518759bc8ull, 0xa4b092bdf66607dfull, 0x91409e3ecac6f383ull, 0x85d0a9c79b1f3919ull,
    0x98a0b5586cf9890full, 0x94f0c0f145e46c85ull, 0x9010cc922b7247f7ull, 0xa210d83b23395debull,
    0x4030e3ec32d3d1a2ull, 0xa5b0efa55fdfa9c4ull, 0xae40fb66affed31aull, 0x8d41073028d7233eull,
    0xa4911301d0125b50ull, 0xa1a11edbab5e2ab5ull, 0xaf712abdc06c31cbull, 0xae8136a814f204aaull,
    0xa661429aaea92ddfull, 0xa9114e95934f312dull, 0x82415a98c8a58e51ull, 0x58f166a45471c3c2ull,
    0xab9172b83c7d517aull, 0x70917ed48695bbc0ull, 0xa7718af9388c8de9ull, 0x94a1972658375d2full,
    0x8e51a35beb6fcb75ull, 0x97b1af99f8138a1cull, 0xa351bbe084045cd3ull, 0x9001c82f95281c6bull,
    0x9e01d4873168b9aaull, 0xa481e0e75eb44026ull, 0xa711ed5022fcd91cull, 0xa201f9c18438ce4cull,
    0x8dc2063b88628cd6ull, 0x935212be3578a819ull, 0x82a21f49917ddc96ull, 0x8d322bdda27912d1ull,
    0x99b2387a6e756238ull, 0x8ac2451ffb82140aull, 0x8ac251ce4fb2a63full, 0x93e25e85711ece75ull,
    0x82b26b4565e27cddull, 0x9e02780e341ddf29ull, 0xa2d284dfe1f56380ull, 0xab4291ba7591bb6full,
    0x86129e9df51fdee1ull, 0xa352ab8a66d10f12ull, 0xafb2b87fd0dad98full, 0xa572c57e39771b2eull,
    0x9002d285a6e4030bull, 0x9d12df961f641589ull, 0x71c2ecafa93e2f56ull, 0xaea2f9d24abd886aull,
    0x86f306fe0a31b715ull, 0x89531432edeeb2fdull, 0x8a932170fc4cd831ull, 0xa1d32eb83ba8ea31ull,
    0x93233c08b26416ffull, 0xab23496266e3fa2cull, 0xa92356c55f929ff0ull, 0xa8f36431a2de883aull,
    0xa4e371a7373aa9caull, 0xa3037f26231e7549ull, 0xa0b38cae6d05d865ull, 0xa3239a401b7140eeull,
    0xad43a7db34e59ff6ull, 0x9543b57fbfec6cf4ull, 0xa083c32dc313a8e4ull, 0x7fe3d0e544ede173ull,
    0x8ad3dea64c123422ull, 0xa943ec70df1c5174ull, 0xa413fa4504ac801bull, 0x8bd40822c367a024ull,
    0xaf04160a21f72e29ull, 0xa3d423fb27094689ull, 0xab8431f5d950a896ull, 0x88843ffa3f84b9d4ull,
    0x48944e086061892dull, 0xae745c2042a7d231ull, 0x9c946a41ed1d0057ull, 0xa1e4786d668b3236ull,
    0x73c486a2b5c13cd0ull, 0xab1494e1e192aed1ull, 0x99c4a32af0d7d3deull, 0xabb4b17dea6db7d6ull,
    0x7d44bfdad5362a27ull, 0x9054ce41b817c114ull, 0x98e4dcb299fddd0dull, 0xa564eb2d81d8abfeull,
    0xa5a4f9b2769d2ca6ull, 0x7a2508417f4531eeull, 0xa82516daa2cf6641ull, 0xac65257de83f4eeeull,
    0xabe5342b569d4f81ull, 0x879542e2f4f6ad27ull, 0xa8a551a4ca5d920eull, 0xa7856070dde910d1ull,
    0x99b56f4736b527daull, 0xa7a57e27dbe2c4ceull, 0x82958d12d497c7fdull, 0xa4059c0827ff07cbull,
    0x9635ab07dd485429ull, 0xa245ba11fba87a02ull, 0x3c45c9268a5946b7ull, 0xa195d84590998b92ull,
    0x9ba5e76f15ad2148ull, 0xa985f6a320dceb70ull, 0xa60605e1b976dc08ull, 0x9e46152ae6cdf6f4ull,
    0xa636247eb03a5584ull, 0x984633dd1d1929fdull, 0xa8e6434634ccc31full, 0xa28652b9febc8fb6ull,
    0xa226623882552224ull, 0xa85671c1c70833f5ull, 0x60368155d44ca973ull, 0x880690f4b19e9538ull,
    0xa216a09e667f3bccull, 0x7a36b052fa75173eull, 0xada6c012750bdabeull, 0x9c76cfdcddd47645ull,
    0xae46dfb23c651a2eull, 0xa7a6ef9298593ae4ull, 0xa9f6ff7df9519483ull, 0x59d70f7466f42e87ull,
    0xaba71f75e8ec5f73ull, 0xa6f72f8286ead089ull, 0xa7a73f9a48a58173ull, 0x90474fbd35d7cbfdull,
    0xa7e75feb564267c8ull, 0x9b777024b1ab6e09ull, 0x986780694fde5d3full, 0x934790b938ac1cf6ull,
    0xaaf7a11473eb0186ull, 0xa207b17b0976cfdaull, 0x9f17c1ed0130c132ull, 0x91b7d26a62ff86f0ull,
    0x7057e2f336cf4e62ull, 0xabe7f3878491c490ull, 0xa6c80427543e1a11ull, 0x946814d2add106d9ull,
    0xa1582589994cce12ull, 0x9998364c1eb941f7ull, 0xa9c8471a4623c7acull, 0xaf2857f4179f5b20ull,
    0xa01868d99b4492ecull, 0x85d879cad931a436ull, 0x99988ac7d98a6699ull, 0x9d589bd0a478580full,
    0x96e8ace5422aa0dbull, 0x9ec8be05bad61778ull, 0xade8cf3216b5448bull, 0xa478e06a5e0866d8ull,
    0x85c8f1ae99157736ull, 0x959902fed0282c8aull, 0xa119145b0b91ffc5ull, 0xab2925c353aa2fe1ull,
    0xae893737b0cdc5e4ull, 0xa88948b82b5f98e4ull, 0xad395a44cbc8520eull, 0xaf296bdd9a7670b2ull,
    0xa1797d829fde4e4full, 0x7ca98f33e47a22a2ull, 0xa749a0f170ca07b9ull, 0xa119b2bb4d53fe0cull,
    0x7c79c49182a3f090ull, 0xa579d674194bb8d4ull, 0x7829e86319e32323ull, 0xaad9fa5e8d07f29dull,
    0xa65a0c667b5de564ull, 0x9c6a1e7aed8eb8bbull, 0x963a309bec4a2d33ull, 0xa2aa42c980460ad7ull,
    0xa16a5503b23e255cull, 0x650a674a8af46052ull, 0x9bca799e1330b358ull, 0xa58a8bfe53c12e58ull,
    0x90fa9e6b5579fdbfull, 0x889ab0e521356ebaull, 0xa81ac36bbfd3f379ull, 0x97ead5ff3a3c2774ull,
    0x97aae89f995ad3adull, 0xa5aafb4ce622f2feull, 0xa21b0e07298db665ull, 0x94db20ce6c9a8952ull,
    0xaedb33a2b84f15faull, 0xac1b468415b749b0ull, 0xa1cb59728de55939ull, 0x92ab6c6e29f1c52aull,
    0xad5b7f76f2fb5e46ull, 0xa24b928cf22749e3ull, 0xa08ba5b030a10649ull, 0xafcbb8e0b79a6f1eull,
    0x823bcc1e904bc1d2ull, 0xafcbdf69c3f3a206ull, 0xa08bf2c25bd71e08ull, 0xa89c06286141b33cull,
    0x811c199bdd85529cull, 0xa48c2d1cd9fa652bull, 0x9b4c40ab5fffd07aull, 0x912c544778fafb22ull,
    0x928c67f12e57d14bull, 0xa86c7ba88988c932ull, 0x71ac8f6d9406e7b5ull, 0xaa0ca3405751c4daull,
    0x750cb720dcef9069ull, 0xac5ccb0f2e6d1674ull, 0xa88cdf0b555dc3f9ull, 0xa2fcf3155b5bab73ull,
    0xa1ad072d4a07897bull, 0x955d1b532b08c968ull, 0xa15d2f87080d89f1ull, 0x93dd43c8eacaa1d6ull,
    0x82ed5818dcfba487ull, 0x5fed6c76e862e6d3ull, 0xa77d80e316c98397ull, 0x9a0d955d71ff6075ull,
    0x9c2da9e603db3285ull, 0xa24dbe7cd63a8314ull, 0x92ddd321f301b460ull, 0xa1ade7d5641c0657ull,
    0xa72dfc97337b9b5eull, 0xadae11676b197d16ull, 0xa42e264614f5a128ull, 0xa30e3b333b16ee11ull,
    0x839e502ee78b3ff6ull, 0xaa7e653924676d75ull, 0x92de7a51fbc74c83ull, 0xa77e8f7977cdb73full,
    0xa0bea4afa2a490d9ull, 0x948eb9f4867cca6eull, 0xa1becf482d8e67f0ull, 0x91cee4aaa2188510ull,
    0x9dcefa1bee615a27ull, 0xa66f0f9c1cb64129ull, 0x93af252b376bba97ull, 0xacdf3ac948dd7273ull,
    0x99df50765b6e4540ull, 0x9faf6632798844f8ull, 0xa12f7bfdad9cbe13ull, 0xaeef91d802243c88ull,
    0x874fa7c1819e90d8ull, 0xacdfbdba3692d513ull, 0x62efd3c22b8f71f1ull, 0x74afe9d96b2a23d9ull,
};

inline double exp(double x) {
    const double MAGIC = 6.755399441055744e15, INV = 369.3299304675746, U = -0.002707606173999011, L = -6.327543041662719e-14;
    if (!(std::fabs(x) <= 708.3964185322641)) return std::exp(x);  // overflow / subnormal results: not needed here
    double N_float = std::fma(x, INV, MAGIC);
    uint64_t nb;
    std::memcpy(&nb, &N_float, 8);
    const int32_t N = (int32_t)(uint32_t)nb;
    N_float -= MAGIC;
    double r = std::fma(N_float, U, x);
    r = std::fma(N_float, L, r);
    const int32_t k = N >> 8;
    const uint64_t j = J_TABLE[N & 255];
    const uint64_t ju = 0x3FF0000000000000ull | (j & 0x000FFFFFFFFFFFFFull), jl = 0x3C00000000000000ull | (j >> 8);
    double jU, jL;
    std::memcpy(&jU, &ju, 8);
    std::memcpy(&jL, &jl, 8);
    const double p = r * std::fma(r, std::fma(r, std::fma(r, 0.04166666857598777, 0.1666666857598779), 0.4999999999999997), 0.9999999999999912);
    const double small_part = std::fma(jU, p, jL) + jU;
    int64_t sb;
    std::memcpy(&sb, &small_part, 8);
    sb += (int64_t)k << 52;
    double out;
    std::memcpy(&out, &sb, 8);
    return out;
}

}  // namespace mbjl
