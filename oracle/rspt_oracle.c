/* TEST INFRASTRUCTURE ONLY -- see rspt_oracle.h.
 *
 * Plain-C restatement of rspt's signal-packer hot path (i_signal_packer compress/decompress for
 * xdelta_hzr, hzr, hadamard, dct).  Written from the behaviour of the reference, stage by stage,
 * in the same decomposition the CUDA pipeline uses (transform -> byte planes -> per-block
 * histogram -> code build -> sized layout -> bit packing), so every kernel has a CPU checker.
 * Citations are relative to /root/reference/lib_rspt/.
 */
#include "rspt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* CRC-32C, reflected polynomial 0x82F63B78, init/xorout ~0 (lib_hzr/hzr_crc32c.c:31-84).       */
/* The table is generated, not transcribed.                                                    */
/* ------------------------------------------------------------------------------------------ */
static uint32_t g_crc_tab[256];
static int g_crc_ready = 0;

static void crc_init(void)
{
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t r = i;
        for (int k = 0; k < 8; ++k)
            r = (r >> 1) ^ (0x82F63B78u & (0u - (r & 1u)));
        g_crc_tab[i] = r;
    }
    g_crc_ready = 1;
}

uint32_t oracle_crc32c(const void* data, size_t n)
{
    if (!g_crc_ready)
        crc_init();
    const uint8_t* p = (const uint8_t*)data;
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i)
        c = (c >> 8) ^ g_crc_tab[(c ^ p[i]) & 0xFFu];
    return ~c;
}

/* ------------------------------------------------------------------------------------------ */
/* hzr: zero-RLE tokens (hzr_internal.h:113-121, hzr_encode.c:133-173 and :410-457)             */
/* ------------------------------------------------------------------------------------------ */
#define NSYM ORACLE_HZR_NSYM
#define BLK ORACLE_HZR_BLOCK
#define RUN_CAP 16662u

/* symbol + extra-bit payload for a zero-run chunk of length z (1..16662) */
static int run_symbol(uint32_t z, uint32_t* extra_val, int* extra_bits)
{
    if (z == 1) { *extra_val = 0; *extra_bits = 0; return 0; }
    if (z == 2) { *extra_val = 0; *extra_bits = 0; return 256; }
    if (z <= 6) { *extra_val = z - 3; *extra_bits = 2; return 257; }
    if (z <= 22) { *extra_val = z - 7; *extra_bits = 4; return 258; }
    if (z <= 278) { *extra_val = z - 23; *extra_bits = 8; return 259; }
    *extra_val = z - 279; *extra_bits = 14; return 260;
}

static const int k_extra_bits[5] = {0, 2, 4, 8, 14}; /* symbols 256..260 */

/* length of the next token's zero-run chunk at in[k] (in[k] == 0): greedy, capped (enc:146-151) */
static uint32_t zero_chunk(const uint8_t* in, size_t n, size_t k)
{
    uint32_t z = 1;
    while (z < RUN_CAP && k + z < n && in[k + z] == 0)
        ++z;
    return z;
}

void oracle_hzr_histogram(const uint8_t* in, size_t n, uint32_t hist[NSYM])
{
    memset(hist, 0, NSYM * sizeof(uint32_t));
    for (size_t k = 0; k < n;) {
        if (in[k] != 0) {
            hist[in[k]]++;
            ++k;
        } else {
            uint32_t z = zero_chunk(in, n, k), ev;
            int eb;
            hist[run_symbol(z, &ev, &eb)]++;
            k += z;
        }
    }
}

/* Huffman tree exactly as MakeTree builds it (hzr_encode.c:222-283): leaves in ascending symbol
 * order; every round joins the two live nodes that are smallest under the total order
 * (count ascending, node index DESCENDING) -- the `<=` at :253/:256 lets the latest index win
 * ties; the smallest becomes child_a (bit 0), the runner-up child_b (bit 1).  Codes are
 * LSB-first (bit `depth` is decided at depth `depth`, :215-218).  The tree is serialised
 * pre-order: branch = 0, leaf = 1 then the 9-bit symbol (:185-189, :209). */
typedef struct {
    uint32_t weight;
    int sym;   /* >= 0 leaf, -1 branch */
    int a, b;  /* children */
} onode;

typedef struct {
    uint8_t* p;
    size_t cap;   /* bytes available */
    size_t nbits; /* bits written */
    int overflow;
} obits;

static void put_bits(obits* w, uint64_t v, int n)
{
    for (int i = 0; i < n; ++i) {
        size_t byte = w->nbits >> 3;
        if (byte >= w->cap) {
            w->overflow = 1;
            return;
        }
        if ((w->nbits & 7) == 0)
            w->p[byte] = 0;
        w->p[byte] |= (uint8_t)(((v >> i) & 1u) << (w->nbits & 7));
        w->nbits++;
    }
}

static void emit_tree(const onode* nd, int root, uint32_t* code, uint8_t* len, obits* w)
{
    /* explicit pre-order stack: (node, code, depth) */
    int st_node[2 * NSYM];
    uint32_t st_code[2 * NSYM];
    int st_depth[2 * NSYM];
    int sp = 0;
    st_node[0] = root; st_code[0] = 0; st_depth[0] = 0; sp = 1;
    while (sp > 0) {
        --sp;
        int k = st_node[sp];
        uint32_t c = st_code[sp];
        int d = st_depth[sp];
        if (nd[k].sym >= 0) {
            put_bits(w, 1, 1);
            put_bits(w, (uint32_t)nd[k].sym, 9);
            code[nd[k].sym] = c;
            len[nd[k].sym] = (uint8_t)d;
        } else {
            put_bits(w, 0, 1);
            /* push b first so that a is visited first */
            st_node[sp] = nd[k].b; st_code[sp] = c | (1u << d); st_depth[sp] = d + 1; ++sp;
            st_node[sp] = nd[k].a; st_code[sp] = c; st_depth[sp] = d + 1; ++sp;
        }
    }
}

int oracle_hzr_build_codes(const uint32_t hist[NSYM], uint32_t code[NSYM], uint8_t len[NSYM],
                           uint8_t* tree, uint32_t* tree_nbits)
{
    onode nd[2 * NSYM];
    int n_leaf = 0;
    memset(code, 0, NSYM * sizeof(uint32_t));
    memset(len, 0, NSYM);
    for (int s = 0; s < NSYM; ++s)
        if (hist[s] > 0) {
            nd[n_leaf].weight = hist[s];
            nd[n_leaf].sym = s;
            nd[n_leaf].a = nd[n_leaf].b = -1;
            ++n_leaf;
        }
    obits w = {tree, 360, 0, 0};
    if (n_leaf == 0) {
        *tree_nbits = 0;
        return 0;
    }
    int n_nodes = n_leaf;
    int live = n_leaf;
    int root = 0;
    while (live > 1) {
        int m1 = -1, m2 = -1; /* smallest, second smallest under (weight asc, index desc) */
        for (int k = 0; k < n_nodes; ++k) {
            if (nd[k].weight == 0)
                continue;
            if (m1 < 0 || nd[k].weight <= nd[m1].weight) {
                m2 = m1;
                m1 = k;
            } else if (m2 < 0 || nd[k].weight <= nd[m2].weight) {
                m2 = k;
            }
        }
        nd[n_nodes].weight = nd[m1].weight + nd[m2].weight;
        nd[n_nodes].sym = -1;
        nd[n_nodes].a = m1;
        nd[n_nodes].b = m2;
        nd[m1].weight = 0;
        nd[m2].weight = 0;
        root = n_nodes++;
        --live;
    }
    if (n_leaf == 1) {
        /* single symbol: a lone leaf stored with a 1-bit code (hzr_encode.c:277-281) */
        put_bits(&w, 1, 1);
        put_bits(&w, (uint32_t)nd[0].sym, 9);
        code[nd[0].sym] = 0;
        len[nd[0].sym] = 1;
    } else {
        emit_tree(nd, root, code, len, &w);
    }
    *tree_nbits = (uint32_t)w.nbits;
    return n_leaf;
}

/* FILL iff every token is of one value class; literal 0 and the run symbols form one class
 * (OnlySingleCode, hzr_encode.c:285-305). */
static int single_class(const uint32_t hist[NSYM])
{
    int zeros = 0, nonzero = 0;
    for (int s = 0; s < NSYM; ++s)
        if (hist[s]) {
            if (s == 0 || s >= 256)
                zeros = 1;
            else
                ++nonzero;
        }
    return (zeros + nonzero) == 1;
}

static uint64_t payload_bits(const uint32_t hist[NSYM], const uint8_t len[NSYM], uint32_t tree_nbits)
{
    uint64_t bits = tree_nbits;
    for (int s = 0; s < NSYM; ++s)
        bits += (uint64_t)hist[s] * (uint64_t)(len[s] + (s >= 256 ? k_extra_bits[s - 256] : 0));
    return bits;
}

int oracle_hzr_block_plan(const uint8_t* in, size_t n, uint32_t* payload_len)
{
    uint32_t hist[NSYM], code[NSYM], tn;
    uint8_t len[NSYM], tree[360];
    oracle_hzr_histogram(in, n, hist);
    if (single_class(hist)) {
        *payload_len = 1;
        return 2;
    }
    oracle_hzr_build_codes(hist, code, len, tree, &tn);
    uint64_t bytes = (payload_bits(hist, len, tn) + 7) >> 3;
    /* the block stream is capped at 7 + n bytes (hzr_encode.c:377-382) and a payload of 65536
     * bytes cannot be described by the 16-bit size field (:466-467) -> plain copy */
    if (bytes > n || bytes >= BLK) {
        *payload_len = (uint32_t)n;
        return 0;
    }
    *payload_len = (uint32_t)bytes;
    return 1;
}

static void put_le16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void put_le32(uint8_t* p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
static uint32_t get_le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static uint32_t get_le32(const uint8_t* p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* One block: 7-byte header (size-1:u16, crc32c(payload):u32, mode:u8) + payload
 * (hzr_encode.c:369-487, hzr_internal.h:84-106).  out must have room for 7 + n bytes. */
static size_t encode_block(const uint8_t* in, size_t n, uint8_t* out)
{
    uint32_t hist[NSYM], code[NSYM], tn = 0;
    uint8_t len[NSYM], tree[360];
    uint8_t* pay = out + 7;
    oracle_hzr_histogram(in, n, hist);
    if (single_class(hist)) {
        pay[0] = in[0];
        put_le16(out, 0);
        put_le32(out + 2, oracle_crc32c(pay, 1));
        out[6] = 2;
        return 8;
    }
    oracle_hzr_build_codes(hist, code, len, tree, &tn);
    uint64_t bytes = (payload_bits(hist, len, tn) + 7) >> 3;
    if (bytes > n || bytes >= BLK) {
        memcpy(pay, in, n);
        put_le16(out, (uint32_t)(n - 1));
        put_le32(out + 2, oracle_crc32c(pay, n));
        out[6] = 0;
        return 7 + n;
    }
    obits w = {pay, n, 0, 0};
    for (uint32_t i = 0; i < tn; ++i)
        put_bits(&w, (tree[i >> 3] >> (i & 7)) & 1u, 1);
    for (size_t k = 0; k < n;) {
        if (in[k] != 0) {
            put_bits(&w, code[in[k]], len[in[k]]);
            ++k;
        } else {
            uint32_t z = zero_chunk(in, n, k), ev;
            int eb;
            int s = run_symbol(z, &ev, &eb);
            put_bits(&w, code[s], len[s]);
            put_bits(&w, ev, eb);
            k += z;
        }
    }
    size_t plen = (w.nbits + 7) >> 3; /* final partial byte is zero padded (enc:84-85) */
    put_le16(out, (uint32_t)(plen - 1));
    put_le32(out + 2, oracle_crc32c(pay, plen));
    out[6] = 1;
    return 7 + plen;
}

size_t oracle_hzr_max_compressed_size(size_t n)
{
    /* hzr_encode.c:489-497 */
    return 4 + (n ? ((n + BLK - 1) / BLK) * 7 + n : 0);
}

int oracle_hzr_encode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* enc)
{
    if (!in || !out || !enc || cap < oracle_hzr_max_compressed_size(n))
        return 1;
    put_le32(out, (uint32_t)n); /* master header: decoded size (enc:521) */
    size_t pos = 4;
    for (size_t off = 0; off < n; off += BLK) {
        size_t m = n - off < BLK ? n - off : BLK;
        pos += encode_block(in + off, m, out + pos);
    }
    *enc = pos;
    return 0;
}

/* ---- decoder (hzr_decode.c:263-333 tree recovery, :335-567 block, :626-674 stream) ---- */
typedef struct {
    const uint8_t* p;
    size_t nbytes;
    size_t pos; /* bit position */
    int fail;
} ibits;

static uint32_t get_bits(ibits* r, int n)
{
    uint32_t v = 0;
    for (int i = 0; i < n; ++i) {
        size_t byte = r->pos >> 3;
        if (byte >= r->nbytes) {
            r->fail = 1;
            return 0;
        }
        v |= (uint32_t)((r->p[byte] >> (r->pos & 7)) & 1u) << i;
        r->pos++;
    }
    return v;
}

typedef struct {
    int sym, a, b;
} dnode;

static int recover_tree(ibits* r, dnode* nd, int* count, int depth)
{
    if (*count >= 2 * NSYM - 1 || depth > 64)
        return -1;
    int me = (*count)++;
    uint32_t leaf = get_bits(r, 1);
    if (r->fail)
        return -1;
    if (leaf) {
        nd[me].sym = (int)get_bits(r, 9);
        nd[me].a = nd[me].b = -1;
        return r->fail ? -1 : me;
    }
    nd[me].sym = -1;
    nd[me].a = recover_tree(r, nd, count, depth + 1);
    if (nd[me].a < 0)
        return -1;
    nd[me].b = recover_tree(r, nd, count, depth + 1);
    if (nd[me].b < 0)
        return -1;
    return me;
}

static int decode_block(const uint8_t* in, size_t avail, uint8_t* out, size_t out_n, size_t* used)
{
    if (avail < 7)
        return 1;
    size_t plen = get_le16(in) + 1;
    int mode = in[6];
    const uint8_t* pay = in + 7;
    if (7 + plen > avail)
        return 1;
    *used = 7 + plen;
    if (mode == 0) {
        if (plen != out_n)
            return 1;
        memcpy(out, pay, out_n);
        return 0;
    }
    if (mode == 2) {
        memset(out, pay[0], out_n);
        return 0;
    }
    if (mode != 1)
        return 1;
    ibits r = {pay, plen, 0, 0};
    dnode nd[2 * NSYM];
    int count = 0;
    int root = recover_tree(&r, nd, &count, 0);
    if (root < 0)
        return 1;
    size_t o = 0;
    while (o < out_n) {
        int k = root;
        if (nd[k].sym >= 0)
            (void)get_bits(&r, 1); /* lone leaf: 1-bit code (dec:487-494) */
        while (nd[k].sym < 0)
            k = get_bits(&r, 1) ? nd[k].b : nd[k].a;
        if (r.fail)
            return 1;
        int s = nd[k].sym;
        if (s <= 255) {
            out[o++] = (uint8_t)s;
        } else {
            size_t z;
            switch (s) {
            case 256: z = 2; break;
            case 257: z = get_bits(&r, 2) + 3; break;
            case 258: z = get_bits(&r, 4) + 7; break;
            case 259: z = get_bits(&r, 8) + 23; break;
            case 260: z = get_bits(&r, 14) + 279; break;
            default: return 1;
            }
            if (r.fail || o + z > out_n)
                return 1;
            memset(out + o, 0, z);
            o += z;
        }
    }
    return 0;
}

int oracle_hzr_decode(const uint8_t* in, size_t n, uint8_t* out, size_t out_size)
{
    if (!in || !out || n < 4)
        return 1;
    size_t total = get_le32(in);
    if (out_size < total)
        return 1;
    size_t pos = 4;
    for (size_t off = 0; off < total; off += BLK) {
        size_t m = total - off < BLK ? total - off : BLK, used = 0;
        if (decode_block(in + pos, n - pos, out + off, m, &used))
            return 1;
        pos += used;
    }
    return pos == n ? 0 : 1;
}

int oracle_hzr_verify(const uint8_t* in, size_t n, size_t* decoded)
{
    /* hzr_decode.c:569-624: walk the block headers and check every payload CRC */
    if (!in || !decoded || n < 4)
        return 1;
    size_t total = get_le32(in), pos = 4;
    *decoded = total;
    for (size_t off = 0; off < total; off += BLK) {
        if (pos + 7 > n)
            return 1;
        size_t plen = get_le16(in + pos) + 1;
        uint32_t want = get_le32(in + pos + 2);
        if (in[pos + 6] > 2 || pos + 7 + plen > n)
            return 1;
        if (oracle_crc32c(in + pos + 7, plen) != want)
            return 1;
        pos += 7 + plen;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* sample conversion and the delta / offset / xor chain (lib_signalpacker/utils.cpp)            */
/* ------------------------------------------------------------------------------------------ */
/* utils.cpp:123-191, little-endian branches: [ns][ch][bps] interleaved -> int32 [ch][ns] with
 * sign extension from 8*bps bits.  (The reference loads an unaligned int32 and shifts; we build
 * the value from the bps bytes so nothing past the buffer is read -- same values.) */
static void native_to_words(const uint8_t* src, int32_t* w, size_t bps, size_t ch, size_t ns)
{
    const int sh = 32 - 8 * (int)bps;
    for (size_t s = 0; s < ns; ++s)
        for (size_t c = 0; c < ch; ++c) {
            const uint8_t* q = src + (s * ch + c) * bps;
            uint32_t v = 0;
            for (size_t b = 0; b < bps; ++b)
                v |= (uint32_t)q[b] << (8 * b);
            w[c * ns + s] = (int32_t)(v << sh) >> sh;
        }
}

/* utils.cpp:51-121: low bps bytes of each word, little-endian, interleaved */
static void words_to_native(const int32_t* w, uint8_t* dst, size_t bps, size_t ch, size_t ns)
{
    for (size_t s = 0; s < ns; ++s)
        for (size_t c = 0; c < ch; ++c) {
            uint32_t v = (uint32_t)w[c * ns + s];
            uint8_t* q = dst + (s * ch + c) * bps;
            for (size_t b = 0; b < bps; ++b)
                q[b] = (uint8_t)(v >> (8 * b));
        }
}

/* delta_encode :193-202, offset_32(-128) :215-219, xor_encode_32 :221-230 over the FLAT array
 * (the chain crosses channel rows).  Unsigned arithmetic = the reference's wrap-around. */
static void xdelta_forward(int32_t* a, size_t n)
{
    uint32_t prev_x = 0, prev_d = 0;
    for (size_t i = 0; i < n; ++i) {
        uint32_t x = (uint32_t)a[i];
        uint32_t d = x - prev_x - 128u;
        a[i] = (int32_t)(d ^ prev_d);
        prev_x = x;
        prev_d = d;
    }
}

/* xor_decode_32 :232-236, offset_32(+128), delta_decode :204-213 */
static void xdelta_inverse(int32_t* a, size_t n)
{
    uint32_t d = 0, x = 0;
    for (size_t i = 0; i < n; ++i) {
        d ^= (uint32_t)a[i];
        x += d + 128u;
        a[i] = (int32_t)x;
    }
}

int32_t oracle_average_32(const int32_t* a, size_t len)
{
    /* utils.cpp:30-40: `int64 sum; sum /= len` with len a size_t => the division is UNSIGNED
     * 64-bit, and the quotient is narrowed to int32. */
    int64_t sum = 0;
    for (size_t i = 0; i < len; ++i)
        sum += a[i];
    return (int32_t)(int64_t)((uint64_t)sum / (uint64_t)len);
}

void oracle_fwht(int n, const int32_t* src, int32_t* dst)
{
    /* fwht.c:4-28: natural-order butterflies, half-span n/2 down to 1, wrap-around int32 */
    uint32_t* a = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    for (int i = 0; i < n; ++i)
        a[i] = (uint32_t)src[i];
    for (int h = n >> 1; h > 0; h >>= 1)
        for (int base = 0; base < n; base += 2 * h)
            for (int j = base; j < base + h; ++j) {
                uint32_t u = a[j], v = a[j + h];
                a[j] = u + v;
                a[j + h] = u - v;
            }
    for (int i = 0; i < n; ++i)
        dst[i] = (int32_t)a[i];
    free(a);
}

/* ------------------------------------------------------------------------------------------ */
/* packers                                                                                      */
/* ------------------------------------------------------------------------------------------ */
struct oracle_packer {
    int kind;
    size_t bps, ch, ns;
    unsigned nb;
    unsigned escalations;
    int32_t* words;   /* [ch*ns] */
    int32_t* tmp;     /* [ns]    */
    uint8_t* planes;  /* [4][ch*ns] */
    uint8_t* verify;  /* [bps*ch*ns] */
    float* cosines;   /* dct: [ns][ns], COS[i][j] = (float)cos((2i+1) j pi / 2n)  (dct.cpp:60-74) */
    double* acc;      /* dct: [ns] */
};

size_t oracle_header_bytes(const oracle_packer* p)
{
    return (p->kind == ORACLE_HADAMARD || p->kind == ORACLE_DCT) ? 3 * p->ch : 0;
}

static int method_byte(int kind)
{
    /* signal_packer_xdelta_hzr.cpp:58, signal_packer_hzr.cpp (0), dct.cpp:127 (1), hadamard.cpp:80 (2) */
    return kind == ORACLE_DCT ? 1 : (kind == ORACLE_HADAMARD ? 2 : 0);
}

oracle_packer* oracle_new(int kind, size_t bps, size_t ch, size_t ns, size_t nb)
{
    if (kind < 0 || kind > 3 || bps < 1 || bps > 4 || !ch || !ns)
        return NULL;
    oracle_packer* p = (oracle_packer*)calloc(1, sizeof(*p));
    p->kind = kind; p->bps = bps; p->ch = ch; p->ns = ns;
    /* planes: xdelta = ctor arg; hzr 4 (hzr.cpp:39); hadamard 3 (:44); dct 2 (:46) */
    p->nb = kind == ORACLE_XDELTA_HZR ? (unsigned)nb : kind == ORACLE_HZR ? 4 : kind == ORACLE_HADAMARD ? 3 : 2;
    if (p->nb < 1 || p->nb > 4) {
        free(p);
        return NULL;
    }
    size_t n = ch * ns;
    p->words = (int32_t*)malloc(n * sizeof(int32_t));
    p->tmp = (int32_t*)malloc(ns * sizeof(int32_t));
    p->planes = (uint8_t*)malloc(4 * n);
    p->verify = (uint8_t*)malloc(bps * n);
    if (kind == ORACLE_DCT) {
        const double PI = 3.14159265358979323846;
        p->cosines = (float*)malloc(ns * ns * sizeof(float));
        p->acc = (double*)malloc(ns * sizeof(double));
        double pi_n_2 = PI / ((double)ns * 2.0);
        for (size_t i = 0; i < ns; ++i)
            for (size_t j = 0; j < ns; ++j)
                p->cosines[i * ns + j] = (float)cos((double)(int)(((int)i << 1) * (int)j + (int)j) * pi_n_2);
    }
    return p;
}

void oracle_delete(oracle_packer* p)
{
    if (!p)
        return;
    free(p->words); free(p->tmp); free(p->planes); free(p->verify); free(p->cosines); free(p->acc);
    free(p);
}

unsigned oracle_nb(const oracle_packer* p) { return p->nb; }
unsigned oracle_escalations(const oracle_packer* p) { return p->escalations; }

size_t oracle_max_compressed_size(const oracle_packer* p)
{
    return 1 + oracle_header_bytes(p) + (size_t)p->nb * (4 + oracle_hzr_max_compressed_size(p->ch * p->ns));
}

/* mean removal shared by hadamard (:59-65) and dct (:104-110); header = low 24 bits, LE (:73-79) */
static void remove_means(oracle_packer* p, uint8_t* header)
{
    for (size_t c = 0; c < p->ch; ++c) {
        int32_t* row = p->words + c * p->ns;
        int32_t m = oracle_average_32(row, p->ns);
        for (size_t i = 0; i < p->ns; ++i)
            row[i] = (int32_t)((uint32_t)row[i] - (uint32_t)m);
        header[3 * c + 0] = (uint8_t)m;
        header[3 * c + 1] = (uint8_t)((uint32_t)m >> 8);
        header[3 * c + 2] = (uint8_t)((uint32_t)m >> 16);
    }
}

static void add_means(oracle_packer* p, const uint8_t* header)
{
    for (size_t c = 0; c < p->ch; ++c) {
        uint32_t v = (uint32_t)header[3 * c] | ((uint32_t)header[3 * c + 1] << 8) | ((uint32_t)header[3 * c + 2] << 16);
        int32_t m = (int32_t)(v << 8) >> 8; /* sign extension from 24 bits (hadamard.cpp:99, dct.cpp:148) */
        int32_t* row = p->words + c * p->ns;
        for (size_t i = 0; i < p->ns; ++i)
            row[i] = (int32_t)((uint32_t)row[i] + (uint32_t)m);
    }
}

/* dct.cpp:76-87.  For every output i: sum(double) += (float)src[x] * COS[x][i] (a FLOAT product)
 * for x ascending; then one multiply by Cs[i]*sqrt(2/n)/128 and truncation toward zero.  The
 * loop nest is interchanged (x outer) for cache friendliness; each sum[i] still receives the
 * same addends in the same order. */
static void dct_forward(oracle_packer* p, const int32_t* src, int32_t* dst)
{
    const size_t n = p->ns;
    const double ratio1 = sqrt(2.0 / (double)(int)n), quality = 128.0;
    const float cs0 = (float)(1 / sqrt(2));
    for (size_t i = 0; i < n; ++i)
        p->acc[i] = 0;
    for (size_t x = 0; x < n; ++x) {
        const float sx = (float)src[x];
        const float* row = p->cosines + x * n;
        for (size_t i = 0; i < n; ++i) {
            float prod = sx * row[i];
            p->acc[i] += prod;
        }
    }
    for (size_t i = 0; i < n; ++i) {
        double sum = p->acc[i];
        sum *= (i ? 1.0f : cs0) * ratio1 / quality;
        dst[i] = (int32_t)sum;
    }
}

/* dct.cpp:89-100: sum += (float)(Cs[x] * (float)dct[x] * COS[i][x]); (int)(sum * sqrt(2/n) * 128) */
static void dct_inverse(oracle_packer* p, const int32_t* coef, int32_t* dst)
{
    const size_t n = p->ns;
    const double ratio1 = sqrt(2.0 / (double)(int)n), quality = 128.0;
    const float cs0 = (float)(1 / sqrt(2));
    for (size_t i = 0; i < n; ++i) {
        double sum = 0;
        const float* row = p->cosines + i * n;
        for (size_t x = 0; x < n; ++x) {
            float t = (x ? 1.0f : cs0) * (float)coef[x];
            float prod = t * row[x];
            sum += prod;
        }
        sum *= ratio1 * quality;
        dst[i] = (int32_t)sum;
    }
}

int oracle_transform(oracle_packer* p, const uint8_t* src, int32_t* words, uint8_t* header)
{
    const size_t n = p->ch * p->ns;
    native_to_words(src, p->words, p->bps, p->ch, p->ns);
    switch (p->kind) {
    case ORACLE_XDELTA_HZR: /* xdelta.cpp:54-57 */
        xdelta_forward(p->words, n);
        break;
    case ORACLE_HZR: /* hzr.cpp: no prediction */
        break;
    case ORACLE_HADAMARD: /* hadamard.cpp:57-72 */
        if (p->ns & (p->ns - 1))
            return 1;
        remove_means(p, header);
        for (size_t c = 0; c < p->ch; ++c) {
            int32_t* row = p->words + c * p->ns;
            oracle_fwht((int)p->ns, row, p->tmp);
            for (size_t i = 0; i < p->ns; ++i) {
                int32_t v = p->tmp[i];
                v /= ((int)p->ns / 1.0); /* fwht_normalize, fwht.c:30-34, ratio = quality = 1 */
                row[i] = v;
            }
        }
        break;
    case ORACLE_DCT: /* dct.cpp:102-119 */
        remove_means(p, header);
        for (size_t c = 0; c < p->ch; ++c) {
            int32_t* row = p->words + c * p->ns;
            dct_forward(p, row, p->tmp);
            memcpy(row, p->tmp, p->ns * sizeof(int32_t));
        }
        xdelta_forward(p->words, n);
        break;
    }
    if (words)
        memcpy(words, p->words, n * sizeof(int32_t));
    return 0;
}

/* p->words holds the reassembled words on entry */
static void inverse_in_place(oracle_packer* p, const uint8_t* header, uint8_t* dst)
{
    const size_t n = p->ch * p->ns;
    switch (p->kind) {
    case ORACLE_XDELTA_HZR: /* xdelta.cpp:80-83 */
        xdelta_inverse(p->words, n);
        break;
    case ORACLE_HZR:
        break;
    case ORACLE_HADAMARD: /* hadamard.cpp:90-101; fwht_normalize2 with ratio 1 is the identity */
        for (size_t c = 0; c < p->ch; ++c) {
            int32_t* row = p->words + c * p->ns;
            oracle_fwht((int)p->ns, row, p->tmp);
            for (size_t i = 0; i < p->ns; ++i) {
                int32_t v = p->tmp[i];
                v /= 1.0;
                row[i] = v;
            }
        }
        add_means(p, header);
        break;
    case ORACLE_DCT: /* dct.cpp:137-150 */
        xdelta_inverse(p->words, n);
        for (size_t c = 0; c < p->ch; ++c) {
            int32_t* row = p->words + c * p->ns;
            dct_inverse(p, row, p->tmp);
            memcpy(row, p->tmp, p->ns * sizeof(int32_t));
        }
        add_means(p, header);
        break;
    }
    words_to_native(p->words, dst, p->bps, p->ch, p->ns);
}

int oracle_inverse(oracle_packer* p, const int32_t* words, const uint8_t* header, uint8_t* dst)
{
    memcpy(p->words, words, p->ch * p->ns * sizeof(int32_t));
    inverse_in_place(p, header, dst);
    return 0;
}

/* compress_i32, signal_packer_base.cpp:38-96: plane k = byte k of every word, flat channel-major;
 * frame = method:u8, header, then per plane len:u32 LE + hzr stream. */
static int frame_encode(oracle_packer* p, const uint8_t* header, uint8_t* dst, size_t cap, size_t* dst_len)
{
    const size_t n = p->ch * p->ns, hb = oracle_header_bytes(p);
    if (cap < oracle_max_compressed_size(p))
        return 1;
    for (unsigned k = 0; k < p->nb; ++k)
        for (size_t i = 0; i < n; ++i)
            p->planes[k * n + i] = (uint8_t)((uint32_t)p->words[i] >> (8 * k));
    size_t pos = 0;
    dst[pos++] = (uint8_t)method_byte(p->kind);
    memcpy(dst + pos, header, hb);
    pos += hb;
    for (unsigned k = 0; k < p->nb; ++k) {
        size_t enc = 0;
        if (oracle_hzr_encode(p->planes + k * n, n, dst + pos + 4, cap - pos - 4, &enc))
            return 1;
        put_le32(dst + pos, (uint32_t)enc);
        pos += 4 + enc;
    }
    *dst_len = pos;
    return 0;
}

/* decompress_i32, signal_packer_base.cpp:98-139: reassemble with sign extension from 8*nb bits */
static int frame_decode(oracle_packer* p, const uint8_t* src, size_t* src_len, uint8_t* header)
{
    const size_t n = p->ch * p->ns, hb = oracle_header_bytes(p);
    int rc = src[0] == method_byte(p->kind) ? 0 : 2; /* "compression method unsupported" */
    size_t pos = 1;
    memcpy(header, src + pos, hb);
    pos += hb;
    memset(p->planes, 0, 4 * n);
    for (unsigned k = 0; k < p->nb; ++k) {
        size_t len = get_le32(src + pos);
        pos += 4;
        if (oracle_hzr_decode(src + pos, len, p->planes + k * n, n))
            rc = rc ? rc : 1;
        pos += len;
    }
    *src_len = pos;
    const int sh = 32 - 8 * (int)p->nb;
    for (size_t i = 0; i < n; ++i) {
        uint32_t v = 0;
        for (unsigned k = 0; k < p->nb; ++k)
            v |= (uint32_t)p->planes[k * n + i] << (8 * k);
        p->words[i] = (int32_t)(v << sh) >> sh;
    }
    return rc;
}

int oracle_decompress(oracle_packer* p, const uint8_t* src, size_t* src_len, uint8_t* dst)
{
    uint8_t header[3 * 256];
    uint8_t* hdr = header;
    uint8_t* big = NULL;
    if (oracle_header_bytes(p) > sizeof(header))
        hdr = big = (uint8_t*)malloc(oracle_header_bytes(p));
    int rc = frame_decode(p, src, src_len, hdr);
    inverse_in_place(p, hdr, dst);
    free(big);
    return rc;
}

int oracle_compress(oracle_packer* p, const uint8_t* src, uint8_t* dst, size_t dst_cap, size_t* dst_len)
{
    uint8_t header[3 * 256];
    uint8_t* hdr = header;
    uint8_t* big = NULL;
    if (oracle_header_bytes(p) > sizeof(header))
        hdr = big = (uint8_t*)malloc(oracle_header_bytes(p));
    int rc = 0;
    for (;;) {
        rc = oracle_transform(p, src, NULL, hdr);
        if (rc)
            break;
        rc = frame_encode(p, hdr, dst, dst_cap, dst_len);
        if (rc || p->kind != ORACLE_XDELTA_HZR)
            break;
        /* xdelta.cpp:59-69: decode what was just written; on mismatch use one more plane, for
         * this and every later frame of this instance. */
        size_t used = 0;
        oracle_decompress(p, dst, &used, p->verify);
        if (memcmp(src, p->verify, p->bps * p->ch * p->ns) == 0)
            break;
        if (p->nb >= 4) {
            rc = 3;
            break;
        }
        p->nb++;
        p->escalations++;
    }
    free(big);
    return rc;
}

size_t oracle_compress_many(oracle_packer* p, const uint8_t* src, size_t frame_bytes, size_t n,
                            uint8_t* dst, size_t dst_stride, uint32_t* sizes)
{
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        size_t len = 0;
        oracle_compress(p, src + i * frame_bytes, dst + i * dst_stride, dst_stride, &len);
        if (sizes)
            sizes[i] = (uint32_t)len;
        total += len;
    }
    return total;
}

size_t oracle_decompress_many(oracle_packer* p, const uint8_t* src, size_t src_stride, size_t n,
                              uint8_t* dst, size_t frame_bytes)
{
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        size_t len = 0;
        oracle_decompress(p, src + i * src_stride, &len, dst + i * frame_bytes);
        total += len;
    }
    return total;
}

double oracle_prdn(const uint8_t* orig, const uint8_t* dec, size_t bps, size_t ch, size_t ns)
{
    /* rspt_test.cpp:98-111: 100*sqrt( sum (o-d)^2 / sum (o-mean_ch)^2 ), mean_ch = average_32 */
    const size_t n = ch * ns;
    int32_t* a = (int32_t*)malloc(n * sizeof(int32_t));
    int32_t* b = (int32_t*)malloc(n * sizeof(int32_t));
    native_to_words(orig, a, bps, ch, ns);
    native_to_words(dec, b, bps, ch, ns);
    double mse = 0, den = 0;
    for (size_t c = 0; c < ch; ++c) {
        int32_t mean = oracle_average_32(a + c * ns, ns);
        for (size_t i = 0; i < ns; ++i) {
            double t = (double)a[c * ns + i] - (double)b[c * ns + i];
            double u = (double)a[c * ns + i] - (double)mean;
            mse += t * t;
            den += u * u;
        }
    }
    free(a);
    free(b);
    return sqrt(mse / den) * 100.0;
}

/* ------------------------------------------------------------------------------------------ */
/* pre-filter (lib_filter/iir_filter.cpp, fir_filter.cpp, as driven by rspt_test.cpp:116-136)  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double x[5], y[5], n[5], d[5];
    int nc;
} iir_state;

static void iir_shift(iir_state* f, double x)
{
    /* iir_filter.cpp:60-65 / :77-82 */
    for (int i = f->nc - 1; i > 0; --i) {
        f->x[i] = f->x[i - 1];
        f->y[i] = f->y[i - 1];
    }
    f->x[0] = x;
}

/* iir_filter::filter, iir_filter.cpp:58-73: the d and n terms alternate */
static double iir_filter_generic(iir_state* f, double x)
{
    iir_shift(f, x);
    double y = f->d[0] * f->x[0];
    for (int i = 1; i < f->nc; ++i) {
        y += f->d[i] * f->x[i];
        y -= f->n[i] * f->y[i];
    }
    f->y[0] = y;
    return y;
}

/* iir_filter::filter_opt, iir_filter.cpp:75-103 with rolling_iir_filter_N_ :26-44: all d terms
 * first, then the n terms, left to right */
static double iir_filter_opt(iir_state* f, double x)
{
    iir_shift(f, x);
    double y = f->d[0] * f->x[0];
    for (int i = 1; i < f->nc; ++i)
        y = y + f->d[i] * f->x[i];
    for (int i = 1; i < f->nc; ++i)
        y = y - f->n[i] * f->y[i];
    f->y[0] = y;
    return y;
}

int oracle_prefilter_iir(uint8_t* frame, size_t bps, size_t ch, size_t ns, const double* n, const double* d,
                         int nr_coefficients, int init_nr_samples)
{
    if (!frame || !n || !d || nr_coefficients < 2 || nr_coefficients > 5 || bps < 1 || bps > 4)
        return 1;
    int32_t* w = (int32_t*)malloc(ch * ns * sizeof(int32_t));
    if (!w)
        return 1;
    native_to_words(frame, w, bps, ch, ns);
    iir_state f;
    memset(&f, 0, sizeof f);
    f.nc = nr_coefficients;
    memcpy(f.n, n, (size_t)nr_coefficients * sizeof(double));
    memcpy(f.d, d, (size_t)nr_coefficients * sizeof(double));
    for (size_t j = 0; j < ch; ++j) {
        int32_t* row = w + j * ns;
        /* init_history_values, iir_filter.cpp:105-109: 4 * nr_samples calls of filter(x); the state
         * of the previous channel is NOT cleared (one object for all channels, rspt_test.cpp:127-133) */
        const double x0 = (double)row[0];
        for (int i = 0; i < 4 * init_nr_samples; ++i)
            iir_filter_generic(&f, x0);
        for (size_t i = 0; i < ns; ++i)
            row[i] = (int32_t)iir_filter_opt(&f, (double)row[i]);   /* rspt_test.cpp:132 */
    }
    words_to_native(w, frame, bps, ch, ns);
    free(w);
    return 0;
}

int oracle_prefilter_fir(uint8_t* frame, size_t bps, size_t ch, size_t ns, const double* kernel, int kernel_size)
{
    if (!frame || !kernel || kernel_size < 1 || bps < 1 || bps > 4)
        return 1;
    const size_t K = (size_t)kernel_size;
    int32_t* w = (int32_t*)malloc(ch * ns * sizeof(int32_t));
    double* ring = (double*)malloc(K * sizeof(double));
    if (!w || !ring) {
        free(w);
        free(ring);
        return 1;
    }
    native_to_words(frame, w, bps, ch, ns);
    for (size_t j = 0; j < ch; ++j) {
        int32_t* row = w + j * ns;
        /* init_history_values, fir_filter.cpp:58-62: kernel_size calls of filter(x) leave the ring
         * holding kernel_size copies of x, whatever it held before (:37-46) */
        for (size_t t = 0; t < K; ++t)
            ring[t] = (double)row[0];
        for (size_t i = 0; i < ns; ++i) {
            /* filter_opt, fir_filter.cpp:48-56: push back, pop front, dot product oldest first */
            memmove(ring, ring + 1, (K - 1) * sizeof(double));
            ring[K - 1] = (double)row[i];
            double y = 0;
            for (size_t t = 0; t < K; ++t)
                y += ring[t] * kernel[t];
            row[i] = (int32_t)y;
        }
    }
    words_to_native(w, frame, bps, ch, ns);
    free(w);
    free(ring);
    return 0;
}
