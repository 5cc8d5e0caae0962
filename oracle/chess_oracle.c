/*
 * chess_oracle.c -- CPU ORACLE (test infrastructure; never linked into the product).
 *
 * Restates, in plain C with ray-walking bitboards, the host-side semantics of the
 * leaf-evaluation path of pierric/smart-chess-rust:
 *
 *   rules      python-chess 1.11.1 (pyproject.toml:10, uv.lock:309-310), reached through
 *              PyO3 at /root/reference/src/chess.rs:665-788 and :356-412.  python-chess is
 *              NOT vendored in the reference tree and is not installable offline, so its
 *              published algorithm (Board.generate_legal_moves / push / is_repetition /
 *              clean_castling_rights / outcome) is restated here FROM MEMORY.
 *              PARITY UNPINNED for: legal-move order beyond the two notebook goldens,
 *              repetition flags, en-passant edge cases, draw claims.  Pinned by: the
 *              notebook move-order goldens (notebooks/verify_model.ipynb:142-161,454-473),
 *              the chess_fast.rs FEN (src/chess_fast.rs:84-98), replay of the 60 SAN games
 *              of py/validation/sample.csv, and perft counts (tests/test_oracle_chess.py).
 *   encode     `_encode`            /root/reference/src/chess.rs:845-877
 *              `BoardHistory::view` /root/reference/src/chess.rs:828-842
 *              `Board::rotate`      /root/reference/src/chess.rs:594-621
 *              `encode_pieces`      /root/reference/src/chess.rs:623-650
 *              `encode_meta`        /root/reference/src/chess.rs:652-662
 *              `Board::extract`     /root/reference/src/chess.rs:356-412
 *   move index `Move::rotate/encode` /root/reference/src/chess.rs:533-550
 *              queenmoves.rs:3-34, knightmoves.rs:7-31, underpromotions.rs:6-33
 *   priors     `post_process_distr` /root/reference/src/chess.rs:879-903
 *   search     `uct/find_max/backward/select/mcts/step` /root/reference/src/mcts.rs:61-328
 *
 * Square numbering is python-chess's: sq = rank*8 + file, a1 = 0, h8 = 63.
 * Colours: WHITE = 1, BLACK = 0 (python-chess chess.WHITE is True; chess.rs:59-63).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>

typedef uint64_t u64;

#define WHITE 1
#define BLACK 0
enum { PAWN = 1, KNIGHT, BISHOP, ROOK, QUEEN, KING };

#define BB(sq) (1ULL << (sq))
#define BB_ALL 0xFFFFFFFFFFFFFFFFULL
#define RANK_BB(r) (0xFFULL << (8 * (r)))
#define FILE_BB(f) (0x0101010101010101ULL << (f))

#define OC_MAX_PLY 1024
#define OC_MAX_MOVES 256

typedef struct {
    u64 pawns, knights, bishops, rooks, queens, kings;
    u64 occ_co[2];
    u64 occupied;
    u64 castling_rights;
    int turn;
    int ep_square; /* -1 = None */
    int halfmove_clock;
    int fullmove_number;
} Pos;

typedef struct {
    uint8_t from, to, promo; /* promo: 0 or piece type */
} OMove;

typedef struct {
    Pos cur;
    int n;                 /* len(move_stack) */
    Pos stack[OC_MAX_PLY]; /* stack[i] = position before move i (python-chess _stack) */
    OMove moves[OC_MAX_PLY];
    int stack_nonempty_at_root; /* 0 for a fresh board: clean_castling_rights recomputes */
} Game;

/* ------------------------------------------------------------------------------------ */
/* tables                                                                               */
/* ------------------------------------------------------------------------------------ */
static u64 T_KNIGHT[64], T_KING[64], T_PAWN[2][64];
static u64 T_RAYS[64][64];
static int g_init = 0;

static inline int msb(u64 b) { return 63 - __builtin_clzll(b); }
static inline int sq_rank(int s) { return s >> 3; }
static inline int sq_file(int s) { return s & 7; }

static u64 step_attacks(int sq, const int (*d)[2], int n)
{
    u64 a = 0;
    int r = sq_rank(sq), f = sq_file(sq);
    for (int i = 0; i < n; i++) {
        int rr = r + d[i][0], ff = f + d[i][1];
        if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) a |= BB(rr * 8 + ff);
    }
    return a;
}

static u64 slide(int sq, u64 occ, const int (*d)[2], int n)
{
    u64 a = 0;
    for (int i = 0; i < n; i++) {
        int r = sq_rank(sq) + d[i][0], f = sq_file(sq) + d[i][1];
        while (r >= 0 && r < 8 && f >= 0 && f < 8) {
            u64 b = BB(r * 8 + f);
            a |= b;
            if (occ & b) break;
            r += d[i][0];
            f += d[i][1];
        }
    }
    return a;
}

static const int D_RANK[2][2] = {{0, 1}, {0, -1}};
static const int D_FILE[2][2] = {{1, 0}, {-1, 0}};
static const int D_DIAG[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};

static inline u64 rank_attacks(int sq, u64 occ) { return slide(sq, occ, D_RANK, 2); }
static inline u64 file_attacks(int sq, u64 occ) { return slide(sq, occ, D_FILE, 2); }
static inline u64 diag_attacks(int sq, u64 occ) { return slide(sq, occ, D_DIAG, 4); }

void oc_init(void)
{
    if (g_init) return;
    static const int KN[8][2] = {{2, 1}, {1, 2}, {-1, 2}, {-2, 1}, {-2, -1}, {-1, -2}, {1, -2}, {2, -1}};
    static const int KG[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
    static const int PW[2][2] = {{1, 1}, {1, -1}};
    static const int PB[2][2] = {{-1, 1}, {-1, -1}};
    for (int s = 0; s < 64; s++) {
        T_KNIGHT[s] = step_attacks(s, KN, 8);
        T_KING[s] = step_attacks(s, KG, 8);
        T_PAWN[WHITE][s] = step_attacks(s, PW, 2);
        T_PAWN[BLACK][s] = step_attacks(s, PB, 2);
    }
    for (int a = 0; a < 64; a++)
        for (int b = 0; b < 64; b++) {
            u64 bb = BB(b), r = 0;
            if (diag_attacks(a, 0) & bb)
                r = (diag_attacks(a, 0) & diag_attacks(b, 0)) | BB(a) | bb;
            else if (rank_attacks(a, 0) & bb)
                r = rank_attacks(a, 0) | BB(a);
            else if (file_attacks(a, 0) & bb)
                r = file_attacks(a, 0) | BB(a);
            T_RAYS[a][b] = r;
        }
    g_init = 1;
}

static inline u64 ray(int a, int b) { return T_RAYS[a][b]; }
static inline u64 between(int a, int b)
{
    u64 bb = T_RAYS[a][b] & ((BB_ALL << a) ^ (BB_ALL << b));
    return bb & (bb - 1);
}

/* ------------------------------------------------------------------------------------ */
/* position primitives                                                                  */
/* ------------------------------------------------------------------------------------ */
static int piece_type_at(const Pos *p, int sq)
{
    u64 m = BB(sq);
    if (!(p->occupied & m)) return 0;
    if (p->pawns & m) return PAWN;
    if (p->knights & m) return KNIGHT;
    if (p->bishops & m) return BISHOP;
    if (p->rooks & m) return ROOK;
    if (p->queens & m) return QUEEN;
    return KING;
}

static int remove_piece_at(Pos *p, int sq)
{
    int pt = piece_type_at(p, sq);
    u64 m = BB(sq);
    if (!pt) return 0;
    p->pawns &= ~m; p->knights &= ~m; p->bishops &= ~m;
    p->rooks &= ~m; p->queens &= ~m; p->kings &= ~m;
    p->occupied ^= m;
    p->occ_co[WHITE] &= ~m;
    p->occ_co[BLACK] &= ~m;
    return pt;
}

static void set_piece_at(Pos *p, int sq, int pt, int color)
{
    remove_piece_at(p, sq);
    u64 m = BB(sq);
    switch (pt) {
    case PAWN: p->pawns |= m; break;
    case KNIGHT: p->knights |= m; break;
    case BISHOP: p->bishops |= m; break;
    case ROOK: p->rooks |= m; break;
    case QUEEN: p->queens |= m; break;
    default: p->kings |= m; break;
    }
    p->occupied ^= m;
    p->occ_co[color] ^= m;
}

static void pos_start(Pos *p)
{
    memset(p, 0, sizeof(*p));
    p->pawns = RANK_BB(1) | RANK_BB(6);
    p->knights = BB(1) | BB(6) | BB(57) | BB(62);
    p->bishops = BB(2) | BB(5) | BB(58) | BB(61);
    p->rooks = BB(0) | BB(7) | BB(56) | BB(63);
    p->queens = BB(3) | BB(59);
    p->kings = BB(4) | BB(60);
    p->occ_co[WHITE] = RANK_BB(0) | RANK_BB(1);
    p->occ_co[BLACK] = RANK_BB(6) | RANK_BB(7);
    p->occupied = p->occ_co[WHITE] | p->occ_co[BLACK];
    p->castling_rights = BB(0) | BB(7) | BB(56) | BB(63);
    p->turn = WHITE;
    p->ep_square = -1;
    p->halfmove_clock = 0;
    p->fullmove_number = 1;
}

static u64 attacks_mask(const Pos *p, int sq)
{
    u64 m = BB(sq);
    if (p->pawns & m) return T_PAWN[(p->occ_co[WHITE] & m) ? WHITE : BLACK][sq];
    if (p->knights & m) return T_KNIGHT[sq];
    if (p->kings & m) return T_KING[sq];
    u64 a = 0;
    if ((p->bishops | p->queens) & m) a = diag_attacks(sq, p->occupied);
    if ((p->rooks | p->queens) & m) a |= rank_attacks(sq, p->occupied) | file_attacks(sq, p->occupied);
    return a;
}

static u64 attackers_mask_occ(const Pos *p, int color, int sq, u64 occ)
{
    u64 qr = p->queens | p->rooks, qb = p->queens | p->bishops;
    u64 a = (T_KING[sq] & p->kings) | (T_KNIGHT[sq] & p->knights) |
            (rank_attacks(sq, occ) & qr) | (file_attacks(sq, occ) & qr) |
            (diag_attacks(sq, occ) & qb) | (T_PAWN[!color][sq] & p->pawns);
    return a & p->occ_co[color];
}

static inline u64 attackers_mask(const Pos *p, int color, int sq)
{
    return attackers_mask_occ(p, color, sq, p->occupied);
}

static int attacked_for_king(const Pos *p, u64 path, u64 occ)
{
    while (path) {
        int s = msb(path);
        path ^= BB(s);
        if (attackers_mask_occ(p, !p->turn, s, occ)) return 1;
    }
    return 0;
}

/* python-chess Board.clean_castling_rights (standard chess).  With a non-empty stack
 * python-chess returns castling_rights unfiltered; push() keeps that field filtered, so
 * recomputing gives the same mask for every position reachable from a legal root. */
static u64 clean_castling_rights(const Pos *p)
{
    u64 c = p->castling_rights & p->rooks;
    u64 w = c & RANK_BB(0) & p->occ_co[WHITE] & (BB(0) | BB(7));
    u64 b = c & RANK_BB(7) & p->occ_co[BLACK] & (BB(56) | BB(63));
    if (!(p->occ_co[WHITE] & p->kings & BB(4))) w = 0;
    if (!(p->occ_co[BLACK] & p->kings & BB(60))) b = 0;
    return w | b;
}

static int has_kingside_castling_rights(const Pos *p, int color)
{
    u64 back = color == WHITE ? RANK_BB(0) : RANK_BB(7);
    u64 king = p->kings & p->occ_co[color] & back;
    if (!king) return 0;
    u64 cr = clean_castling_rights(p) & back;
    while (cr) {
        u64 rook = cr & -cr;
        if (rook > king) return 1;
        cr &= cr - 1;
    }
    return 0;
}

static int has_queenside_castling_rights(const Pos *p, int color)
{
    u64 back = color == WHITE ? RANK_BB(0) : RANK_BB(7);
    u64 king = p->kings & p->occ_co[color] & back;
    if (!king) return 0;
    u64 cr = clean_castling_rights(p) & back;
    while (cr) {
        u64 rook = cr & -cr;
        if (rook < king) return 1;
        cr &= cr - 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* move generation, python-chess order                                                  */
/* ------------------------------------------------------------------------------------ */
typedef struct { OMove m[OC_MAX_MOVES]; int n; } MoveList;

static inline void ml_add(MoveList *l, int from, int to, int promo)
{
    l->m[l->n].from = (uint8_t)from;
    l->m[l->n].to = (uint8_t)to;
    l->m[l->n].promo = (uint8_t)promo;
    l->n++;
}

static void ml_add_pawn(MoveList *l, int from, int to)
{
    int r = sq_rank(to);
    if (r == 0 || r == 7) {
        ml_add(l, from, to, QUEEN);
        ml_add(l, from, to, ROOK);
        ml_add(l, from, to, BISHOP);
        ml_add(l, from, to, KNIGHT);
    } else
        ml_add(l, from, to, 0);
}

static void gen_pseudo_legal_ep(const Pos *p, u64 from_mask, u64 to_mask, MoveList *l)
{
    if (p->ep_square < 0 || !(BB(p->ep_square) & to_mask)) return;
    if (BB(p->ep_square) & p->occupied) return;
    u64 capturers = p->pawns & p->occ_co[p->turn] & from_mask & T_PAWN[!p->turn][p->ep_square] &
                    RANK_BB(p->turn == WHITE ? 4 : 3);
    while (capturers) {
        int c = msb(capturers);
        capturers ^= BB(c);
        ml_add(l, c, p->ep_square, 0);
    }
}

static void gen_castling(const Pos *p, u64 from_mask, u64 to_mask, MoveList *l)
{
    u64 back = p->turn == WHITE ? RANK_BB(0) : RANK_BB(7);
    u64 king = p->occ_co[p->turn] & p->kings & back & from_mask;
    king &= -king;
    if (!king) return;
    u64 bb_c = FILE_BB(2) & back, bb_d = FILE_BB(3) & back, bb_f = FILE_BB(5) & back, bb_g = FILE_BB(6) & back;
    u64 cands = clean_castling_rights(p) & back & to_mask;
    while (cands) {
        int cand = msb(cands);
        cands ^= BB(cand);
        u64 rook = BB(cand);
        int a_side = rook < king;
        u64 king_to = a_side ? bb_c : bb_g;
        u64 rook_to = a_side ? bb_d : bb_f;
        u64 king_path = between(msb(king), msb(king_to));
        u64 rook_path = between(cand, msb(rook_to));
        if (!(((p->occupied ^ king ^ rook) & (king_path | rook_path | king_to | rook_to)) ||
              attacked_for_king(p, king_path | king, p->occupied ^ king) ||
              attacked_for_king(p, king_to, p->occupied ^ king ^ rook ^ rook_to))) {
            /* standard chess: reported as the king's two-square move */
            ml_add(l, msb(king), msb(king_to), 0);
        }
    }
}

static void gen_pseudo_legal(const Pos *p, u64 from_mask, u64 to_mask, MoveList *l)
{
    u64 ours = p->occ_co[p->turn];
    u64 non_pawns = ours & ~p->pawns & from_mask;
    while (non_pawns) {
        int from = msb(non_pawns);
        non_pawns ^= BB(from);
        u64 mv = attacks_mask(p, from) & ~ours & to_mask;
        while (mv) {
            int to = msb(mv);
            mv ^= BB(to);
            ml_add(l, from, to, 0);
        }
    }
    if (from_mask & p->kings) gen_castling(p, from_mask, to_mask, l);

    u64 pawns = p->pawns & ours & from_mask;
    if (!pawns) return;

    u64 cap = pawns;
    while (cap) {
        int from = msb(cap);
        cap ^= BB(from);
        u64 targets = T_PAWN[p->turn][from] & p->occ_co[!p->turn] & to_mask;
        while (targets) {
            int to = msb(targets);
            targets ^= BB(to);
            ml_add_pawn(l, from, to);
        }
    }
    u64 single, dbl;
    if (p->turn == WHITE) {
        single = (pawns << 8) & ~p->occupied;
        dbl = (single << 8) & ~p->occupied & (RANK_BB(2) | RANK_BB(3));
    } else {
        single = (pawns >> 8) & ~p->occupied;
        dbl = (single >> 8) & ~p->occupied & (RANK_BB(5) | RANK_BB(4));
    }
    single &= to_mask;
    dbl &= to_mask;
    while (single) {
        int to = msb(single);
        single ^= BB(to);
        ml_add_pawn(l, to + (p->turn == BLACK ? 8 : -8), to);
    }
    while (dbl) {
        int to = msb(dbl);
        dbl ^= BB(to);
        ml_add(l, to + (p->turn == BLACK ? 16 : -16), to, 0);
    }
    if (p->ep_square >= 0) gen_pseudo_legal_ep(p, from_mask, to_mask, l);
}

static u64 slider_blockers(const Pos *p, int king)
{
    u64 rq = p->rooks | p->queens, bq = p->bishops | p->queens;
    u64 snipers = (rank_attacks(king, 0) & rq) | (file_attacks(king, 0) & rq) | (diag_attacks(king, 0) & bq);
    u64 blockers = 0;
    snipers &= p->occ_co[!p->turn];
    while (snipers) {
        int s = msb(snipers);
        snipers ^= BB(s);
        u64 b = between(king, s) & p->occupied;
        if (b && BB(msb(b)) == b) blockers |= b;
    }
    return blockers & p->occ_co[p->turn];
}

static int is_en_passant(const Pos *p, const OMove *m)
{
    int d = (int)m->to - (int)m->from;
    if (d < 0) d = -d;
    return p->ep_square == m->to && (p->pawns & BB(m->from)) && (d == 7 || d == 9) &&
           !(p->occupied & BB(m->to));
}

static int is_castling(const Pos *p, const OMove *m)
{
    if (p->kings & BB(m->from)) {
        int d = sq_file(m->from) - sq_file(m->to);
        if (d < 0) d = -d;
        return d > 1 || ((p->rooks & p->occ_co[p->turn] & BB(m->to)) != 0);
    }
    return 0;
}

static u64 pin_mask(const Pos *p, int color, int sq)
{
    u64 kbb = p->kings & p->occ_co[color];
    if (!kbb) return BB_ALL;
    int king = msb(kbb);
    u64 sm = BB(sq);
    u64 rq = p->rooks | p->queens, bq = p->bishops | p->queens;
    u64 rays3[3] = {file_attacks(king, 0), rank_attacks(king, 0), diag_attacks(king, 0)};
    u64 sl3[3] = {rq, rq, bq};
    for (int i = 0; i < 3; i++) {
        if (rays3[i] & sm) {
            u64 snipers = rays3[i] & sl3[i] & p->occ_co[!color];
            while (snipers) {
                int s = msb(snipers);
                snipers ^= BB(s);
                if ((between(s, king) & (p->occupied | sm)) == sm) return ray(king, s);
            }
            break;
        }
    }
    return BB_ALL;
}

static int ep_skewered(const Pos *p, int king, int capturer)
{
    int last_double = p->ep_square + (p->turn == WHITE ? -8 : 8);
    u64 occ = (p->occupied & ~BB(last_double) & ~BB(capturer)) | BB(p->ep_square);
    u64 ha = p->occ_co[!p->turn] & (p->rooks | p->queens);
    if (rank_attacks(king, occ) & ha) return 1;
    u64 da = p->occ_co[!p->turn] & (p->bishops | p->queens);
    if (diag_attacks(king, occ) & da) return 1;
    return 0;
}

static int is_safe(const Pos *p, int king, u64 blockers, const OMove *m)
{
    if (m->from == king) {
        if (is_castling(p, m)) return 1;
        return !attackers_mask(p, !p->turn, m->to);
    } else if (is_en_passant(p, m)) {
        return (pin_mask(p, p->turn, m->from) & BB(m->to)) && !ep_skewered(p, king, m->from);
    } else {
        return !(blockers & BB(m->from)) || (ray(m->from, m->to) & BB(king));
    }
}

static void gen_evasions(const Pos *p, int king, u64 checkers, u64 from_mask, u64 to_mask, MoveList *l)
{
    u64 sliders = checkers & (p->bishops | p->rooks | p->queens);
    u64 attacked = 0;
    while (sliders) {
        int c = msb(sliders);
        sliders ^= BB(c);
        attacked |= ray(king, c) & ~BB(c);
    }
    if (BB(king) & from_mask) {
        u64 mv = T_KING[king] & ~p->occ_co[p->turn] & ~attacked & to_mask;
        while (mv) {
            int to = msb(mv);
            mv ^= BB(to);
            ml_add(l, king, to, 0);
        }
    }
    int checker = msb(checkers);
    if (BB(checker) == checkers) {
        u64 target = between(king, checker) | checkers;
        gen_pseudo_legal(p, ~p->kings & from_mask, target & to_mask, l);
        if (p->ep_square >= 0 && !(BB(p->ep_square) & target)) {
            int last_double = p->ep_square + (p->turn == WHITE ? -8 : 8);
            if (last_double == checker) gen_pseudo_legal_ep(p, from_mask, to_mask, l);
        }
    }
}

static void gen_legal_masked(const Pos *p, u64 from_mask, u64 to_mask, MoveList *out)
{
    MoveList tmp;
    tmp.n = 0;
    out->n = 0;
    u64 kbb = p->kings & p->occ_co[p->turn];
    if (!kbb) {
        gen_pseudo_legal(p, from_mask, to_mask, out);
        return;
    }
    int king = msb(kbb);
    u64 blockers = slider_blockers(p, king);
    u64 checkers = attackers_mask(p, !p->turn, king);
    if (checkers)
        gen_evasions(p, king, checkers, from_mask, to_mask, &tmp);
    else
        gen_pseudo_legal(p, from_mask, to_mask, &tmp);
    for (int i = 0; i < tmp.n; i++)
        if (is_safe(p, king, blockers, &tmp.m[i])) out->m[out->n++] = tmp.m[i];
}

static void gen_legal(const Pos *p, MoveList *out) { gen_legal_masked(p, BB_ALL, BB_ALL, out); }

static int has_legal_en_passant(const Pos *p)
{
    if (p->ep_square < 0) return 0;
    /* generate_legal_ep: legal moves restricted to ep captures */
    MoveList tmp, l;
    tmp.n = 0;
    gen_pseudo_legal_ep(p, BB_ALL, BB_ALL, &tmp);
    if (!tmp.n) return 0;
    gen_legal(p, &l);
    for (int i = 0; i < l.n; i++)
        if (is_en_passant(p, &l.m[i])) return 1;
    return 0;
}

static int is_check(const Pos *p)
{
    u64 kbb = p->kings & p->occ_co[p->turn];
    if (!kbb) return 0;
    return attackers_mask(p, !p->turn, msb(kbb)) != 0;
}

/* ------------------------------------------------------------------------------------ */
/* push (python-chess Board.push, standard chess)                                       */
/* ------------------------------------------------------------------------------------ */
static int is_zeroing(const Pos *p, const OMove *m)
{
    u64 touched = BB(m->from) ^ BB(m->to);
    return (touched & p->pawns) || (touched & p->occ_co[!p->turn]);
}

static void pos_push(Pos *p, const OMove *m)
{
    p->castling_rights = clean_castling_rights(p);
    int ep_square = p->ep_square;
    p->ep_square = -1;
    p->halfmove_clock += 1;
    if (p->turn == BLACK) p->fullmove_number += 1;
    if (is_zeroing(p, m)) p->halfmove_clock = 0;

    u64 from_bb = BB(m->from), to_bb = BB(m->to);
    int pt = remove_piece_at(p, m->from);
    int capture_square = m->to;
    int captured = piece_type_at(p, capture_square);

    p->castling_rights &= ~to_bb & ~from_bb;
    if (pt == KING) {
        if (p->turn == WHITE) p->castling_rights &= ~RANK_BB(0);
        else p->castling_rights &= ~RANK_BB(7);
    } else if (captured == KING) {
        if (p->turn == WHITE && sq_rank(m->to) == 7) p->castling_rights &= ~RANK_BB(7);
        else if (p->turn == BLACK && sq_rank(m->to) == 0) p->castling_rights &= ~RANK_BB(0);
    }

    if (pt == PAWN) {
        int diff = (int)m->to - (int)m->from;
        if (diff == 16 && sq_rank(m->from) == 1) p->ep_square = m->from + 8;
        else if (diff == -16 && sq_rank(m->from) == 6) p->ep_square = m->from - 8;
        else if (m->to == ep_square && (diff == 7 || diff == 9 || diff == -7 || diff == -9) && !captured) {
            int down = p->turn == WHITE ? -8 : 8;
            capture_square = ep_square + down;
            captured = remove_piece_at(p, capture_square);
        }
    }
    if (m->promo) pt = m->promo;

    /* castling arrives as the king's two-square move; python-chess converts it to
     * king-takes-own-rook (_to_chess960) before this point */
    int castling = 0;
    if (pt == KING) {
        int df = sq_file(m->to) - sq_file(m->from);
        if (df == 2 || df == -2) {
            int a_side = df < 0;
            int rook_from = a_side ? (p->turn == WHITE ? 0 : 56) : (p->turn == WHITE ? 7 : 63);
            if (p->rooks & p->occ_co[p->turn] & BB(rook_from)) {
                castling = 1;
                /* rights bookkeeping on the rook square as python-chess does with to_bb */
                p->castling_rights &= ~BB(rook_from);
                remove_piece_at(p, rook_from);
                if (a_side) {
                    set_piece_at(p, p->turn == WHITE ? 2 : 58, KING, p->turn);
                    set_piece_at(p, p->turn == WHITE ? 3 : 59, ROOK, p->turn);
                } else {
                    set_piece_at(p, p->turn == WHITE ? 6 : 62, KING, p->turn);
                    set_piece_at(p, p->turn == WHITE ? 5 : 61, ROOK, p->turn);
                }
            }
        }
    }
    if (!castling) set_piece_at(p, m->to, pt, p->turn);
    p->turn = !p->turn;
}

/* ------------------------------------------------------------------------------------ */
/* Game: position + stack                                                               */
/* ------------------------------------------------------------------------------------ */
Game *oc_game_new(void)
{
    oc_init();
    Game *g = (Game *)calloc(1, sizeof(Game));
    pos_start(&g->cur);
    return g;
}

void oc_game_free(Game *g) { free(g); }

void oc_game_copy(Game *dst, const Game *src)
{
    dst->cur = src->cur;
    dst->n = src->n;
    memcpy(dst->stack, src->stack, sizeof(Pos) * (size_t)src->n);
    memcpy(dst->moves, src->moves, sizeof(OMove) * (size_t)src->n);
}

Game *oc_game_dup(const Game *g)
{
    Game *d = (Game *)malloc(sizeof(Game));
    oc_game_copy(d, g);
    return d;
}

static int parse_fen_board(Pos *p, const char *fen)
{
    memset(p, 0, sizeof(*p));
    int r = 7, f = 0;
    const char *c = fen;
    for (; *c && *c != ' '; c++) {
        if (*c == '/') { r--; f = 0; continue; }
        if (*c >= '1' && *c <= '8') { f += *c - '0'; continue; }
        int color = (*c >= 'A' && *c <= 'Z') ? WHITE : BLACK;
        int pt;
        switch (*c | 0x20) {
        case 'p': pt = PAWN; break;
        case 'n': pt = KNIGHT; break;
        case 'b': pt = BISHOP; break;
        case 'r': pt = ROOK; break;
        case 'q': pt = QUEEN; break;
        case 'k': pt = KING; break;
        default: return -1;
        }
        if (r < 0 || f > 7) return -1;
        set_piece_at(p, r * 8 + f, pt, color);
        f++;
    }
    if (*c != ' ') return -1;
    c++;
    p->turn = (*c == 'w') ? WHITE : BLACK;
    c += 2;
    p->castling_rights = 0;
    for (; *c && *c != ' '; c++) {
        if (*c == 'K') p->castling_rights |= BB(7);
        if (*c == 'Q') p->castling_rights |= BB(0);
        if (*c == 'k') p->castling_rights |= BB(63);
        if (*c == 'q') p->castling_rights |= BB(56);
    }
    if (*c) c++;
    p->ep_square = -1;
    if (*c && *c != '-') {
        p->ep_square = (c[1] - '1') * 8 + (c[0] - 'a');
        c += 2;
    } else if (*c)
        c++;
    p->halfmove_clock = 0;
    p->fullmove_number = 1;
    if (*c == ' ') {
        int h = 0, fm = 1;
        if (sscanf(c, " %d %d", &h, &fm) >= 1) {
            p->halfmove_clock = h;
            p->fullmove_number = fm < 1 ? 1 : fm;
        }
    }
    return 0;
}

Game *oc_game_from_fen(const char *fen)
{
    oc_init();
    Game *g = (Game *)calloc(1, sizeof(Game));
    if (parse_fen_board(&g->cur, fen)) { free(g); return NULL; }
    return g;
}

int oc_game_ply(const Game *g) { return g->n; }
int oc_game_turn(const Game *g) { return g->cur.turn; }

int oc_game_push(Game *g, int from, int to, int promo)
{
    if (g->n >= OC_MAX_PLY) return -1;
    OMove m = {(uint8_t)from, (uint8_t)to, (uint8_t)promo};
    g->stack[g->n] = g->cur;
    g->moves[g->n] = m;
    g->n++;
    pos_push(&g->cur, &m);
    return 0;
}

int oc_game_pop(Game *g)
{
    if (g->n == 0) return -1;
    g->n--;
    g->cur = g->stack[g->n];
    return 0;
}

/* out: n x 3 bytes (from, to, promo); returns n */
int oc_game_legal_moves(const Game *g, uint8_t *out)
{
    MoveList l;
    gen_legal(&g->cur, &l);
    for (int i = 0; i < l.n; i++) {
        out[3 * i] = l.m[i].from;
        out[3 * i + 1] = l.m[i].to;
        out[3 * i + 2] = l.m[i].promo;
    }
    return l.n;
}

int oc_game_is_check(const Game *g) { return is_check(&g->cur); }
int oc_game_piece_at(const Game *g, int sq)
{
    int pt = piece_type_at(&g->cur, sq);
    if (!pt) return 0;
    return (g->cur.occ_co[WHITE] & BB(sq)) ? pt : -pt;
}

/* position at ply k of the game (k = g->n is the current one) */
static const Pos *pos_at(const Game *g, int k) { return k == g->n ? &g->cur : &g->stack[k]; }

typedef struct {
    u64 pawns, knights, bishops, rooks, queens, kings, w, b, cr;
    int turn, ep;
} TKey;

static void transposition_key(const Pos *p, TKey *k)
{
    k->pawns = p->pawns; k->knights = p->knights; k->bishops = p->bishops;
    k->rooks = p->rooks; k->queens = p->queens; k->kings = p->kings;
    k->w = p->occ_co[WHITE]; k->b = p->occ_co[BLACK];
    k->turn = p->turn;
    k->cr = clean_castling_rights(p);
    k->ep = has_legal_en_passant(p) ? p->ep_square : -1;
}

static int tkey_eq(const TKey *a, const TKey *b)
{
    return a->pawns == b->pawns && a->knights == b->knights && a->bishops == b->bishops &&
           a->rooks == b->rooks && a->queens == b->queens && a->kings == b->kings && a->w == b->w &&
           a->b == b->b && a->turn == b->turn && a->cr == b->cr && a->ep == b->ep;
}

static int reduces_castling_rights(const Pos *p, const OMove *m)
{
    u64 cr = clean_castling_rights(p);
    u64 touched = BB(m->from) ^ BB(m->to);
    /* castling is stored as the king move; python-chess evaluates this on the
     * king-takes-rook form, both forms touch the king square */
    return (touched & cr) || ((cr & RANK_BB(0)) && (touched & p->kings & p->occ_co[WHITE])) ||
           ((cr & RANK_BB(7)) && (touched & p->kings & p->occ_co[BLACK]));
}

static int is_irreversible(const Pos *p, const OMove *m)
{
    return is_zeroing(p, m) || reduces_castling_rights(p, m) || has_legal_en_passant(p);
}

/* python-chess Board.is_repetition(count) evaluated on the position at ply `at`
 * with the move stack moves[0..at). */
static int is_repetition_at(const Game *g, int at, int count)
{
    const Pos *cur = pos_at(g, at);
    int maybe = 1;
    for (int k = at - 1; k >= 0; k--) {
        if (g->stack[k].occupied == cur->occupied) {
            maybe++;
            if (maybe >= count) break;
        }
    }
    if (maybe < count) return 0;

    TKey key, k2;
    transposition_key(cur, &key);
    int len = at; /* len(move_stack) while popping */
    for (;;) {
        if (count <= 1) return 1;
        if (len < count - 1) break;
        /* move = self.pop() */
        len--;
        const Pos *prev = &g->stack[len];
        const OMove *mv = &g->moves[len];
        if (is_irreversible(prev, mv)) break;
        transposition_key(prev, &k2);
        if (tkey_eq(&k2, &key)) count--;
    }
    return 0;
}

int oc_game_is_repetition(const Game *g, int count) { return is_repetition_at(g, g->n, count); }

/* ------------------------------------------------------------------------------------ */
/* encode: `_encode` (chess.rs:845-877)                                                 */
/* ------------------------------------------------------------------------------------ */
/* planes: int8 [8][8][112] (rank, file, channel) HWC as `Array3<i8>`; meta int32[7].
 * node_depth = depth of the tree node below the game root (history stops at the root:
 * n_hist = min(8, node_depth + 1), chess.rs:851-867). */
void oc_game_encode(const Game *g, int node_depth, int8_t *planes, int32_t *meta)
{
    memset(planes, 0, 8 * 8 * 112);
    const Pos *cur = &g->cur;
    int rotate = cur->turn == BLACK; /* node.step.1 == state.turn() (torch.rs:111) */
    int n_hist = node_depth + 1;
    if (n_hist > 8) n_hist = 8;
    if (n_hist > g->n + 1) n_hist = g->n + 1;
    for (int t = 0; t < n_hist; t++) {
        int ply = g->n - t;
        const Pos *p = pos_at(g, ply);
        int rep2 = is_repetition_at(g, ply, 2);
        int rep3 = is_repetition_at(g, ply, 3);
        for (int sq = 0; sq < 64; sq++) {
            int pt = piece_type_at(p, sq);
            if (!pt) continue;
            int color = (p->occ_co[WHITE] & BB(sq)) ? WHITE : BLACK;
            int rank = sq_rank(sq), file = sq_file(sq);
            if (rotate) { /* Board::rotate: rank -> 7-rank, colours swapped */
                rank = 7 - rank;
                color = !color;
            }
            int ch = 14 * t + (pt - 1) + (color == WHITE ? 0 : 6);
            planes[(rank * 8 + file) * 112 + ch] = 1;
        }
        if (rep2 || rep3)
            for (int s = 0; s < 64; s++) {
                if (rep2) planes[s * 112 + 14 * t + 12] = 1;
                if (rep3) planes[s * 112 + 14 * t + 13] = 1;
            }
    }
    /* encode_meta on the UNROTATED current board (chess.rs:874) */
    meta[0] = cur->turn;
    meta[1] = cur->fullmove_number;
    meta[2] = has_kingside_castling_rights(cur, cur->turn);
    meta[3] = has_queenside_castling_rights(cur, cur->turn);
    meta[4] = has_kingside_castling_rights(cur, !cur->turn);
    meta[5] = has_queenside_castling_rights(cur, !cur->turn);
    meta[6] = cur->halfmove_clock;
}

/* What the Rust shim hands to the C ABI (include/sc_b200.h, sc_position): the inputs of
 * `_encode` BEFORE rotation/encoding -- per history slot the python-chess bitboards and the
 * two repetition flags of `Board::extract` (chess.rs:375-380), plus `encode_meta` of the
 * current board.  slot layout: [P,N,B,R,Q,K, white occupancy, flags]. */
void oc_game_pack(const Game *g, int node_depth, u64 *slot /*[8][8]*/, int32_t *meta, int32_t *n_hist_out)
{
    memset(slot, 0, sizeof(u64) * 64);
    int n_hist = node_depth + 1;
    if (n_hist > 8) n_hist = 8;
    if (n_hist > g->n + 1) n_hist = g->n + 1;
    for (int t = 0; t < n_hist; t++) {
        int ply = g->n - t;
        const Pos *p = pos_at(g, ply);
        u64 *s = slot + 8 * t;
        s[0] = p->pawns; s[1] = p->knights; s[2] = p->bishops;
        s[3] = p->rooks; s[4] = p->queens; s[5] = p->kings;
        s[6] = p->occ_co[WHITE];
        s[7] = (u64)(is_repetition_at(g, ply, 2) ? 1 : 0) | (u64)(is_repetition_at(g, ply, 3) ? 2 : 0);
    }
    const Pos *cur = &g->cur;
    meta[0] = cur->turn;
    meta[1] = cur->fullmove_number;
    meta[2] = has_kingside_castling_rights(cur, cur->turn);
    meta[3] = has_queenside_castling_rights(cur, cur->turn);
    meta[4] = has_kingside_castling_rights(cur, !cur->turn);
    meta[5] = has_queenside_castling_rights(cur, !cur->turn);
    meta[6] = cur->halfmove_clock;
    *n_hist_out = n_hist;
}

/* ------------------------------------------------------------------------------------ */
/* move index (queenmoves.rs / knightmoves.rs / underpromotions.rs)                     */
/* ------------------------------------------------------------------------------------ */
int oc_move_index(int from, int to, int promo, int turn)
{
    int fr = sq_rank(from), ff = sq_file(from), tr = sq_rank(to), tf = sq_file(to);
    if (turn == BLACK) { fr = 7 - fr; tr = 7 - tr; } /* Move::rotate */
    int d0 = tr - fr, d1 = tf - ff;
    int a0 = d0 < 0 ? -d0 : d0, a1 = d1 < 0 ? -d1 : d1;
    int queen_promo = (promo == 0) || (promo == QUEEN);
    if ((d0 == 0 || d1 == 0 || a0 == a1) && queen_promo) {
        int dist = a0 > a1 ? a0 : a1;
        int s0 = (d0 > 0) - (d0 < 0), s1 = (d1 > 0) - (d1 < 0);
        int dir;
        if (s0 == -1 && s1 == -1) dir = 5;
        else if (s0 == -1 && s1 == 0) dir = 4;
        else if (s0 == -1 && s1 == 1) dir = 3;
        else if (s0 == 0 && s1 == -1) dir = 6;
        else if (s0 == 0 && s1 == 1) dir = 2;
        else if (s0 == 1 && s1 == -1) dir = 7;
        else if (s0 == 1 && s1 == 0) dir = 0;
        else if (s0 == 1 && s1 == 1) dir = 1;
        else return -1;
        return fr * 8 * 73 + ff * 73 + dir * 7 + (dist - 1);
    }
    static const int KD[8][2] = {{2, 1}, {1, 2}, {-1, 2}, {-2, 1}, {-2, -1}, {-1, -2}, {1, -2}, {2, -1}};
    for (int k = 0; k < 8; k++)
        if (KD[k][0] == d0 && KD[k][1] == d1) return fr * 8 * 73 + ff * 73 + 56 + k;
    if ((promo == KNIGHT || promo == BISHOP || promo == ROOK) && fr == 6 && tr == 7) {
        if (d1 < -1 || d1 > 1) return -1;
        int pi = promo == KNIGHT ? 0 : (promo == BISHOP ? 1 : 2);
        return fr * 8 * 73 + ff * 73 + 64 + (d1 + 1) * 3 + pi;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------ */
/* outcome(claim_draw=True) -- python-chess Board.outcome (chess.rs:719-729)            */
/* termination codes follow chess.rs:87-105; winner: 1 white, 0 black, -1 none          */
/* ------------------------------------------------------------------------------------ */
static int has_insufficient_material(const Pos *p, int color)
{
    u64 own = p->occ_co[color];
    if (own & (p->pawns | p->rooks | p->queens)) return 0;
    if (own & p->knights) {
        int cnt = __builtin_popcountll(own);
        return cnt <= 2 && !(p->occ_co[!color] & ~p->kings & ~p->queens);
    }
    if (own & p->bishops) {
        const u64 DARK = 0xAA55AA55AA55AA55ULL, LIGHT = 0x55AA55AA55AA55AAULL;
        int same_color = !(p->bishops & DARK) || !(p->bishops & LIGHT);
        return same_color && !p->pawns && !p->knights;
    }
    return 1;
}

static int is_repetition_game(Game *g, int count) { return is_repetition_at(g, g->n, count); }

int oc_game_outcome(Game *g, int claim_draw, int *winner)
{
    MoveList l;
    gen_legal(&g->cur, &l);
    *winner = -1;
    if (l.n == 0 && is_check(&g->cur)) { *winner = !g->cur.turn; return 1; }
    if (has_insufficient_material(&g->cur, WHITE) && has_insufficient_material(&g->cur, BLACK)) return 3;
    if (l.n == 0) return 2;
    if (g->cur.halfmove_clock >= 150) return 4;
    if (is_repetition_game(g, 5)) return 5;
    if (claim_draw) {
        /* can_claim_fifty_moves: is_fifty_moves(), or clock >= 99 and some non-zeroing
         * legal move leads to a position with clock >= 100 that still has a legal move */
        if (g->cur.halfmove_clock >= 100) return 6;
        if (g->cur.halfmove_clock >= 99) {
            for (int i = 0; i < l.n; i++)
                if (!is_zeroing(&g->cur, &l.m[i])) {
                    oc_game_push(g, l.m[i].from, l.m[i].to, l.m[i].promo);
                    MoveList l2;
                    gen_legal(&g->cur, &l2);
                    oc_game_pop(g);
                    if (l2.n > 0) return 6;
                }
        }
        /* can_claim_threefold_repetition: count transposition keys back to the last
         * irreversible move; claim if the current key occurred 3x, or if some legal
         * move reaches a key already seen 2x */
        static TKey keys[OC_MAX_PLY + 1];
        int nk = 0;
        transposition_key(&g->cur, &keys[nk++]);
        for (int len = g->n; len > 0;) {
            len--;
            if (is_irreversible(&g->stack[len], &g->moves[len])) break;
            transposition_key(&g->stack[len], &keys[nk++]);
        }
        int cnt = 0;
        for (int k = 0; k < nk; k++) cnt += tkey_eq(&keys[k], &keys[0]);
        if (cnt >= 3) return 7;
        for (int i = 0; i < l.n; i++) {
            TKey nkey;
            oc_game_push(g, l.m[i].from, l.m[i].to, l.m[i].promo);
            transposition_key(&g->cur, &nkey);
            oc_game_pop(g);
            int c2 = 0;
            for (int k = 0; k < nk; k++) c2 += tkey_eq(&keys[k], &nkey);
            if (c2 >= 2) return 7;
        }
    }
    return 0;
}

/* perft for pinning the generator against published node counts */
static u64 perft_rec(Game *g, int depth)
{
    MoveList l;
    gen_legal(&g->cur, &l);
    if (depth == 1) return (u64)l.n;
    u64 n = 0;
    for (int i = 0; i < l.n; i++) {
        oc_game_push(g, l.m[i].from, l.m[i].to, l.m[i].promo);
        n += perft_rec(g, depth - 1);
        oc_game_pop(g);
    }
    return n;
}

u64 oc_perft(Game *g, int depth) { return depth <= 0 ? 1 : perft_rec(g, depth); }

/* ------------------------------------------------------------------------------------ */
/* priors: post_process_distr (chess.rs:879-903), argmax = false                        */
/* ------------------------------------------------------------------------------------ */
void oc_post_process(const float *logp, const int32_t *idx, int n, float *out)
{
    float sum = 0.f;
    for (int i = 0; i < n; i++) {
        out[i] = expf(logp[idx[i]]);
        sum += out[i];
    }
    sum += 1e-5f;
    for (int i = 0; i < n; i++) out[i] = out[i] / sum;
}

/* ------------------------------------------------------------------------------------ */
/* sequential PUCT search (mcts.rs)                                                     */
/* ------------------------------------------------------------------------------------ */
/* evaluator callback = `Game::predict` for a non-terminal position: fills priors[n_moves]
 * (already post-processed) and returns the value in White's perspective. */
typedef float (*oc_eval_fn)(void *ctx, const Game *g, int node_depth, const uint8_t *moves, int n_moves,
                            float *priors);

typedef struct ONode {
    OMove mv;
    int step_color; /* Step.1: side to move AFTER the move (chess.rs:65-66) */
    int depth;
    float q;
    int n;
    float uct;
    struct ONode *parent;
    struct ONode *children;
    int n_children;
    /* cache of predict() at this node: mcts.rs:152 re-evaluates the node on every
     * descent; predict is a pure function of (node, state), so caching is equivalent */
    float *prior;
    float value;
    int evaluated; /* 0 no, 1 yes, 2 terminal */
} ONode;

typedef struct {
    ONode *root;      /* game root (depth 0) */
    ONode *cursor;    /* current search root */
    Game *game;       /* state at cursor */
    oc_eval_fn eval;
    void *ctx;
    long n_evals;     /* evaluator calls (cached mode) */
    long n_predicts;  /* predict() calls the reference would have made (F2) */
} OTree;

OTree *oc_tree_new(oc_eval_fn eval, void *ctx)
{
    OTree *t = (OTree *)calloc(1, sizeof(OTree));
    t->root = (ONode *)calloc(1, sizeof(ONode));
    t->root->step_color = WHITE; /* Step(None, White) main.rs:157-165 */
    t->cursor = t->root;
    t->game = oc_game_new();
    t->eval = eval;
    t->ctx = ctx;
    return t;
}

static void node_free_children(ONode *n)
{
    for (int i = 0; i < n->n_children; i++) {
        node_free_children(&n->children[i]);
        free(n->children[i].prior);
    }
    free(n->children);
    n->children = NULL;
    n->n_children = 0;
}

void oc_tree_free(OTree *t)
{
    node_free_children(t->root);
    free(t->root->prior);
    free(t->root);
    oc_game_free(t->game);
    free(t);
}

Game *oc_tree_game(OTree *t) { return t->game; }
long oc_tree_n_evals(const OTree *t) { return t->n_evals; }
long oc_tree_n_predicts(const OTree *t) { return t->n_predicts; }

/* predict(): legal moves, terminal value or evaluator (torch.rs:89-146) */
static int node_predict(OTree *t, ONode *node, Game *state, uint8_t *moves, float **prior, float *value)
{
    int n = oc_game_legal_moves(state, moves);
    t->n_predicts++;
    if (n == 0) {
        int w;
        oc_game_outcome(state, 1, &w);
        *value = w == WHITE ? 1.f : (w == BLACK ? -1.f : 0.f);
        *prior = NULL;
        return 0;
    }
    if (!node->evaluated) {
        node->prior = (float *)malloc(sizeof(float) * (size_t)n);
        node->value = t->eval(t->ctx, state, node->depth, moves, n, node->prior);
        node->evaluated = 1;
        t->n_evals++;
    }
    *prior = node->prior;
    *value = node->value;
    return n;
}

static float uct_score(float sqrt_total, float prior, float q, int n_act, int reverse_q, float cpuct)
{
    float avg = q / ((float)n_act + 1e-4f) * (reverse_q ? -1.f : 1.f);
    float expl = (sqrt_total + 0.01f) / (1.f + (float)n_act) * cpuct * prior;
    return avg + expl;
}

/* one rollout of mcts.rs:261-288 (with_noise = false) */
static void rollout(OTree *t, Game *state, float cpuct)
{
    ONode *path[OC_MAX_PLY];
    int plen = 0;
    uint8_t moves[3 * OC_MAX_MOVES];
    float *prior, value;
    ONode *node = t->cursor;
    path[plen++] = node;
    int n_steps;
    for (;;) {
        n_steps = node_predict(t, node, state, moves, &prior, &value);
        int reverse_q = node->step_color == BLACK;
        if (node->n_children == 0 || n_steps == 0) break;
        ONode *best;
        if (node->n_children == 1)
            best = &node->children[0];
        else {
            int tot = 0;
            for (int i = 0; i < node->n_children; i++) tot += node->children[i].n;
            float sq = sqrtf((float)tot);
            int bi = 0;
            float bu = 0.f;
            for (int i = 0; i < node->n_children; i++) {
                ONode *c = &node->children[i];
                float u = uct_score(sq, prior[i], c->q, c->n, reverse_q, cpuct);
                c->uct = u;
                if (i == 0 || u >= bu) { bu = u; bi = i; } /* max_by: last max wins */
            }
            best = &node->children[bi];
        }
        oc_game_push(state, best->mv.from, best->mv.to, best->mv.promo);
        path[plen++] = best;
        node = best;
    }
    /* expansion (mcts.rs:269-283): children replaced by the returned steps */
    if (node->n_children) node_free_children(node);
    if (n_steps > 0) {
        node->children = (ONode *)calloc((size_t)n_steps, sizeof(ONode));
        node->n_children = n_steps;
        for (int i = 0; i < n_steps; i++) {
            ONode *c = &node->children[i];
            c->mv.from = moves[3 * i];
            c->mv.to = moves[3 * i + 1];
            c->mv.promo = moves[3 * i + 2];
            c->step_color = !state->cur.turn;
            c->depth = node->depth + 1;
            c->parent = node;
        }
    }
    for (int i = 0; i < plen; i++) {
        path[i]->n += 1;
        path[i]->q += value;
    }
}

void oc_tree_search(OTree *t, int n_rollout, float cpuct)
{
    Game *local = (Game *)malloc(sizeof(Game));
    for (int r = 0; r < n_rollout; r++) {
        oc_game_copy(local, t->game);
        rollout(t, local, cpuct);
    }
    free(local);
}

int oc_tree_root_children(const OTree *t, uint8_t *moves, int32_t *n_act, float *q, float *uct)
{
    const ONode *c = t->cursor;
    for (int i = 0; i < c->n_children; i++) {
        moves[3 * i] = c->children[i].mv.from;
        moves[3 * i + 1] = c->children[i].mv.to;
        moves[3 * i + 2] = c->children[i].mv.promo;
        n_act[i] = c->children[i].n;
        q[i] = c->children[i].q;
        uct[i] = c->children[i].uct;
    }
    return c->n_children;
}

float oc_tree_root_q(const OTree *t) { return t->cursor->q; }
int oc_tree_root_n(const OTree *t) { return t->cursor->n; }

/* mcts::step with temperature 0 (mcts.rs:309-311): FIRST child with the max visit count;
 * then navigate down and reset the chosen child (mcts.rs:319-323). Returns child index or -1. */
int oc_tree_step_argmax(OTree *t)
{
    ONode *c = t->cursor;
    if (c->n_children == 0) return -1;
    int best = 0;
    for (int i = 1; i < c->n_children; i++)
        if (c->children[i].n > c->children[best].n) best = i;
    ONode *ch = &c->children[best];
    t->cursor = ch;
    ch->q = 0.f;
    ch->n = 0;
    node_free_children(ch);
    /* the cached evaluation of the node is a pure function of the position: keep it */
    oc_game_push(t->game, ch->mv.from, ch->mv.to, ch->mv.promo);
    return best;
}

int oc_tree_step_index(OTree *t, int idx)
{
    ONode *c = t->cursor;
    if (idx < 0 || idx >= c->n_children) return -1;
    ONode *ch = &c->children[idx];
    t->cursor = ch;
    ch->q = 0.f;
    ch->n = 0;
    node_free_children(ch);
    oc_game_push(t->game, ch->mv.from, ch->mv.to, ch->mv.promo);
    return idx;
}

/* ------------------------------------------------------------------------------------ */
/* deterministic stand-in evaluator shared (by specification, not by code) with the      */
/* product's test hook: priors/value from an integer hash of the position.  Lets the     */
/* search semantics be compared without NN numerics in the loop.                         */
/* ------------------------------------------------------------------------------------ */
static inline u64 mix64(u64 x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

u64 oc_position_hash(const Game *g)
{
    const Pos *p = &g->cur;
    u64 h = mix64(p->pawns);
    h = mix64(h ^ p->knights); h = mix64(h ^ p->bishops); h = mix64(h ^ p->rooks);
    h = mix64(h ^ p->queens); h = mix64(h ^ p->kings); h = mix64(h ^ p->occ_co[WHITE]);
    h = mix64(h ^ (u64)p->turn);
    h = mix64(h ^ clean_castling_rights(p));
    h = mix64(h ^ (u64)(p->ep_square + 1));
    return h;
}

float oc_hash_eval(void *ctx, const Game *g, int node_depth, const uint8_t *moves, int n, float *priors)
{
    (void)ctx; (void)node_depth;
    u64 h = oc_position_hash(g);
    float sum = 0.f;
    for (int i = 0; i < n; i++) {
        u64 m = mix64(h ^ ((u64)moves[3 * i] << 16) ^ ((u64)moves[3 * i + 1] << 8) ^ moves[3 * i + 2]);
        priors[i] = (float)((m >> 40) + 1) * (1.0f / 16777216.0f);
        sum += priors[i];
    }
    sum += 1e-5f;
    for (int i = 0; i < n; i++) priors[i] = priors[i] / sum;
    u64 v = mix64(h ^ 0xABCDEF);
    return ((float)(v >> 40) * (1.0f / 16777216.0f)) * 2.f - 1.f;
}

OTree *oc_tree_new_hash(void) { return oc_tree_new(oc_hash_eval, NULL); }
