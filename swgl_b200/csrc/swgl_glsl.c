/*
 * swgl_glsl.c -- quirk-compatible front-end for the reference's GLSL subset (plain C, host).
 *
 * Stage 1 scans the source the way the reference tokenizer does (swgl.c:852-1867) and builds
 * a small AST; stage 2 type-checks it with the interpreter's rules (swgl.c:2177-2840) and
 * emits the straight-line IR of swgl_ir.h.  Nothing here runs per vertex or per fragment.
 *
 * Where the reference would crash or hang on malformed input (NULL token dereference,
 * cursor that stops advancing) this front-end reports a compile error instead and the
 * shader is marked not-ok; a program using it draws nothing.
 */
#include "swgl_glsl.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* "not found" of the scanners.  Far beyond any source (swgl_glsl_compile refuses longer ones) and small enough that the
 * `cursor = position + 1` of a caller that has not looked at it yet cannot overflow: the cursor then lies behind the end
 * of the text, every scanner loop is empty and ch() reads as 0. */
#define BIG 0x3FFFFFFF

/* ------------------------------------------------------------------------------------------
 * literals (swgl.c:18-88)
 * ---------------------------------------------------------------------------------------- */
double swgl_glsl_atof(const char* s)
{
	double value = 0.0, scale = 1.0, sign = 1.0;
	int seen_point = 0;
	if (*s == '-') { sign = -1.0; s++; }
	for (; *s; s++)
	{
		if (*s == '.') { seen_point = 1; continue; }
		int d = *s - '0';
		if (d < 0 || d > 9) break;
		if (seen_point) { scale /= 10.0; value += scale * d; }
		else value = value * 10.0 + d;
	}
	return value * sign;
}

int swgl_glsl_atoi(const char* s)
{
	int value = 0, sign = 1;
	while (*s == ' ') s++;
	if (*s == '-' || *s == '+') { sign = (*s == '-') ? -1 : 1; s++; }
	while (*s >= '0' && *s <= '9') { value = (int)((unsigned)value * 10u + (unsigned)(*s - '0')); s++; }
	return sign * value;
}

/* ------------------------------------------------------------------------------------------
 * AST
 * ---------------------------------------------------------------------------------------- */
enum
{
	N_VAR, N_CONST, N_BIN, N_ASSIGN, N_DECL, N_SWZ, N_CALL, N_CONS
};
enum { OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_LT, OP_GT, OP_EQ, OP_ASSIGN };
enum { FN_TEXTURE, FN_COS, FN_SIN, FN_TAN, FN_MIN, FN_MAX };

typedef struct node
{
	int kind;
	int op;                 /* N_BIN: OP_*; N_CALL: FN_*; N_CONS: SWT_* */
	struct node* a;
	struct node* b;
	struct node* args[8];
	int n_args;
	int bad_arg;            /* an argument failed to parse */
	int var;
	int is_float; float f; int i;
	int swz[8]; int n_swz;
} node;

#define MAX_NODES 2048
#define MAX_LINES 256

typedef struct
{
	char* code;
	int size;
	int at;
	swgl_shader* sh;
	int scope_first, scope_last;    /* locals of the function being parsed: vars[first..last) */
	node nodes[MAX_NODES];
	int n_nodes;
	node* main_lines[MAX_LINES];
	int n_main_lines;
	int failed;
} parser;

static void fail(parser* p, const char* msg)
{
	if (!p->failed)
	{
		p->failed = 1;
		snprintf(p->sh->error, sizeof(p->sh->error), "%s (near offset %d)", msg, p->at);
	}
}

static node* new_node(parser* p, int kind)
{
	if (p->n_nodes >= MAX_NODES) { fail(p, "shader too large"); return NULL; }
	node* n = &p->nodes[p->n_nodes++];
	memset(n, 0, sizeof(*n));
	n->kind = kind;
	n->var = -1;
	return n;
}

static char ch(const parser* p, int i)
{
	return (i >= 0 && i < p->size) ? p->code[i] : 0;
}

/* ---- scanners (swgl.c:852-898) ---- */
static int tell_next(const parser* p, char c)
{
	for (int i = p->at; i < p->size; i++) if (p->code[i] == c) return i;
	return BIG;
}

static int tell_next_matching(const parser* p, char inc, char dec)
{
	int depth = 0;
	for (int i = p->at; i < p->size; i++)
	{
		if (p->code[i] == inc) depth++;
		if (p->code[i] == dec) { depth--; if (depth == 0) return i; }
	}
	return BIG;
}

static int tell_next_arg_start(const parser* p)
{
	int depth = 0;
	for (int i = p->at; i < p->size; i++)
	{
		if (p->code[i] == '(') depth++;
		if (p->code[i] == ')') depth--;
		if (depth == 0 && p->code[i] == ',') return i;
	}
	return BIG;
}

static int is_op_char(char c)
{
	return c == '+' || c == '-' || c == '*' || c == '/' || c == '<' || c == '>' || c == '=';
}
static int is_digit_or_minus(char c) { return (c >= '0' && c <= '9') || c == '-'; } /* swgl.c:444 */

/* swgl.c:910-988.  Returns the operator span [first, second) and its kind, or first = BIG. */
static void tell_next_operator(const parser* p, int* first, int* second, int* op)
{
	int depth = 0;
	for (int i = p->at; i < p->size; i++)
	{
		char c = p->code[i];
		if (depth == 0 && is_op_char(c))
		{
			char n = ch(p, i + 1);
			int two = is_op_char(n);
			if (!two)
			{
				int k = -1;
				if (c == '+') k = OP_ADD;
				else if (c == '-') { if (!is_digit_or_minus(n)) k = OP_SUB; }
				else if (c == '*') k = OP_MUL;
				else if (c == '/') k = OP_DIV;
				else if (c == '=') k = OP_ASSIGN;
				else if (c == '<') k = OP_LT;
				else if (c == '>') k = OP_GT;
				if (k >= 0) { *first = i; *second = i + 1; *op = k; return; }
			}
			else if (c == '=' && n == '=')
			{
				*first = i; *second = i + 2; *op = OP_EQ; return;
			}
			/* any other two-character operator spelling matches nothing here */
		}
		if (c == '(') depth++;
		if (c == ')') depth--;
	}
	*first = BIG; *second = BIG;
}

/* copy [at, idx) (optionally stopping at the first blank) into out */
static void str_until(const parser* p, int idx, int stop_at_blank, char* out, int out_len)
{
	if (idx == BIG) { snprintf(out, out_len, "ERROR"); return; }
	int n = 0;
	for (int i = p->at; i < idx && n < out_len - 1; i++)
	{
		char c = ch(p, i);
		if (stop_at_blank && c == ' ') break;
		out[n++] = c;
	}
	out[n] = 0;
}

static int type_from_str(const char* s) /* swgl.c:791-803 */
{
	if (!strcmp(s, "vec2")) return SWT_VEC2;
	if (!strcmp(s, "vec3")) return SWT_VEC3;
	if (!strcmp(s, "vec4")) return SWT_VEC4;
	if (!strcmp(s, "float")) return SWT_FLOAT;
	if (!strcmp(s, "int")) return SWT_INT;
	if (!strcmp(s, "mat2")) return SWT_MAT2;
	if (!strcmp(s, "mat3")) return SWT_MAT3;
	if (!strcmp(s, "mat4")) return SWT_MAT4;
	if (!strcmp(s, "sampler2D")) return SWT_SAMPLER2D;
	return SWT_UNKNOWN;
}

static void skip_blanks(parser* p) { while (ch(p, p->at) == ' ') p->at++; }

/* ---- variables ---- */
static int add_var(parser* p, const char* name, int type, int is_local)
{
	swgl_shader* sh = p->sh;
	if (sh->n_vars >= SWGL_MAX_VARS) { fail(p, "too many variables"); return -1; }
	uint32_t word = 0;
	if (sh->n_vars > 0)
	{
		const swgl_var* last = &sh->vars[sh->n_vars - 1];
		word = last->word + (uint32_t)swt_words(last->type);
	}
	if (word + (uint32_t)swt_words(type) > SWGL_MAX_VAR_WORDS) { fail(p, "variable file too large"); return -1; }
	swgl_var* v = &sh->vars[sh->n_vars];
	memset(v, 0, sizeof(*v));
	snprintf(v->name, sizeof(v->name), "%s", name);
	v->type = type;
	v->is_local = (uint8_t)is_local;
	v->word = word;
	v->copy_of = -1;
	v->location = -1;
	if (!is_local) sh->n_globals = sh->n_vars + 1;
	return sh->n_vars++;
}

/* GLSLFindVariable (swgl.c:1004-1017): globals first, then the function scope. */
static int find_var(const parser* p, const char* name)
{
	const swgl_shader* sh = p->sh;
	for (int i = 0; i < sh->n_vars; i++)
		if (!sh->vars[i].is_local && !strcmp(sh->vars[i].name, name)) return i;
	for (int i = p->scope_first; i < p->scope_last; i++)
		if (sh->vars[i].is_local && !strcmp(sh->vars[i].name, name)) return i;
	return -1;
}

/* ---- expressions ---- */
static node* parse_expr(parser* p, int end_at);

/* swizzle letters after a '.', up to end_at (swgl.c:1110-1142, 1230-1262, 1330-1358) */
static int parse_swizzle(parser* p, node* sw, int from, int end_at, int strict)
{
	for (int i = from; i < end_at; i++)
	{
		char c = ch(p, i);
		int pick = -1;
		if (c == 'x' || c == 's') pick = 0;
		else if (c == 'y' || c == 't') pick = 1;
		else if (c == 'z') pick = 2;
		else if (c == 'w') pick = 3;
		else if (c == ' ') break;
		else if (strict) return 0;
		else continue;
		if (sw->n_swz < 8) sw->swz[sw->n_swz] = pick;
		sw->n_swz++;
	}
	return 1;
}

static node* parse_args(parser* p, int end_at) /* swgl.c:1045-1071 */
{
	node* call = new_node(p, N_CALL);
	if (!call) return NULL;
	int guard = 0;
	while (p->at < end_at)
	{
		if (++guard > 64) { fail(p, "argument list does not terminate"); return NULL; }
		skip_blanks(p);
		int next = tell_next_arg_start(p);
		if (next > end_at) next = end_at;
		node* arg = parse_expr(p, next);
		if (!arg) call->bad_arg = 1;
		if (call->n_args < 8) call->args[call->n_args] = arg;
		call->n_args++;
		if (next == end_at) break;
		p->at = next + 1;
		skip_blanks(p);
	}
	p->at = end_at + 1;
	return call;
}

static node* parse_subexpr(parser* p, int end_at) /* swgl.c:1074-1370 */
{
	skip_blanks(p);
	/* swizzle detection stops at the first blank -- "(a + b).xy" therefore has none */
	int depth = 0, swz = -1;
	for (int i = p->at; i < end_at; i++)
	{
		char c = ch(p, i);
		if (c == ' ') break;
		if (c == '(') depth++;
		if (c == ')') depth--;
		if (c == '.' && depth == 0) { swz = i + 1; break; }
	}

	if (ch(p, p->at) == '(')
	{
		int close = tell_next_matching(p, '(', ')');
		if (close == BIG) { fail(p, "unbalanced parenthesis"); return NULL; }
		p->at++;
		node* inner = parse_expr(p, close);
		if (!inner) { fail(p, "empty parenthesis"); return NULL; }
		if (swz != -1)
		{
			node* sw = new_node(p, N_SWZ);
			if (!sw) return NULL;
			sw->a = inner;
			if (!parse_swizzle(p, sw, swz, end_at, 1)) { fail(p, "bad swizzle"); return NULL; }
			inner = sw;
		}
		p->at = end_at + 1;
		return inner;
	}

	int open = tell_next(p, '(');
	char word[SWGL_NAME_LEN * 2];
	str_until(p, open, 1, word, sizeof(word));
	int ctor = type_from_str(word);
	if (ctor != SWT_UNKNOWN)
	{
		p->at = open;
		int close = tell_next_matching(p, '(', ')');
		if (close == BIG) { fail(p, "unbalanced constructor"); return NULL; }
		p->at++;
		node* call = parse_args(p, close);
		if (!call) return NULL;
		int want = ctor == SWT_FLOAT ? 1 : ctor == SWT_VEC2 ? 2 : ctor == SWT_VEC3 ? 3
		         : ctor == SWT_VEC4 ? 4 : ctor == SWT_INT ? 1 : -1;
		if (want < 0) { fail(p, "matrix/sampler constructors are not executable in the reference"); return NULL; }
		if (call->n_args != want || call->bad_arg) { fail(p, "constructor needs exactly one scalar per component"); return NULL; }
		call->kind = N_CONS;
		call->op = ctor;
		p->at = end_at + 1;
		return call;
	}

	int fn = -1;
	if (!strcmp(word, "texture")) fn = FN_TEXTURE;
	else if (!strcmp(word, "cos")) fn = FN_COS;
	else if (!strcmp(word, "sin")) fn = FN_SIN;
	else if (!strcmp(word, "tan")) fn = FN_TAN;
	else if (!strcmp(word, "min")) fn = FN_MIN;
	else if (!strcmp(word, "max")) fn = FN_MAX;
	if (fn >= 0)
	{
		p->at = open;
		int close = tell_next_matching(p, '(', ')');
		if (close == BIG) { fail(p, "unbalanced call"); return NULL; }
		p->at++;
		node* call = parse_args(p, close);
		if (!call) return NULL;
		if (call->bad_arg) { fail(p, "empty call argument"); return NULL; }
		call->op = fn;
		node* out = call;
		if (swz != -1)
		{
			node* sw = new_node(p, N_SWZ);
			if (!sw) return NULL;
			sw->a = call;
			if (!parse_swizzle(p, sw, swz, end_at, 1)) { fail(p, "bad swizzle"); return NULL; }
			out = sw;
		}
		p->at = end_at + 1;
		return out;
	}

	if (is_digit_or_minus(ch(p, p->at)))
	{
		node* k = new_node(p, N_CONST);
		if (!k) return NULL;
		int is_float = 0;
		for (int i = p->at; i < end_at; i++) if (ch(p, i) == '.') { is_float = 1; break; }
		char num[64];
		str_until(p, end_at, 1, num, sizeof(num));
		k->is_float = is_float;
		if (is_float) k->f = (float)swgl_glsl_atof(num); /* double -> float, as `Fval = swgl_atof()` */
		else k->i = swgl_glsl_atoi(num);
		p->at = end_at + 1;
		return k;
	}

	/* variable, optionally swizzled */
	char name[SWGL_NAME_LEN * 2];
	int n = 0, vswz = -1;
	for (int i = p->at; i < end_at; i++)
	{
		char c = ch(p, i);
		if (c == ' ') break;
		if (c == '.') { vswz = i + 1; break; }
		if (n < (int)sizeof(name) - 1) name[n++] = c;
	}
	name[n] = 0;
	int var = find_var(p, name);
	if (var < 0)
	{
		char msg[160];
		snprintf(msg, sizeof(msg), "unknown identifier '%.60s'", name);
		fail(p, msg);
		return NULL;
	}
	node* v = new_node(p, N_VAR);
	if (!v) return NULL;
	v->var = var;
	node* out = v;
	if (vswz != -1)
	{
		node* sw = new_node(p, N_SWZ);
		if (!sw) return NULL;
		sw->a = v;
		parse_swizzle(p, sw, vswz, end_at, 0); /* unknown letters are skipped here (swgl.c:1330-1358) */
		out = sw;
	}
	p->at = end_at + 1;
	return out;
}

/* swgl.c:1372-1408: sub (op sub)* folded strictly left to right, no precedence */
static node* parse_expr(parser* p, int end_at)
{
	skip_blanks(p);
	if (p->at == end_at) return NULL;
	if (p->at > end_at) return NULL;

	int first, second, op = -1;
	tell_next_operator(p, &first, &second, &op);
	node* tok = parse_subexpr(p, first < end_at ? first : end_at);
	if (!tok) return NULL;

	int guard = 0;
	while (p->at < end_at && first != BIG)
	{
		if (++guard > 256) { fail(p, "expression does not terminate"); return NULL; }
		node* bin = new_node(p, op == OP_ASSIGN ? N_ASSIGN : N_BIN);
		if (!bin) return NULL;
		bin->op = op;
		bin->a = tok;
		p->at = second;
		tell_next_operator(p, &first, &second, &op);
		bin->b = parse_subexpr(p, first < end_at ? first : end_at);
		if (!bin->b) { fail(p, "operator without right operand"); return NULL; }
		tok = bin;
	}
	p->at = end_at + 1;
	return tok;
}

/* swgl.c:1410-1488.  Returns NULL for "no executable token" (uninitialised declaration). */
static node* parse_line(parser* p)
{
	skip_blanks(p);
	int blank = tell_next(p, ' ');
	int semi = tell_next(p, ';');
	int first, second, op = -1;
	tell_next_operator(p, &first, &second, &op);
	if (semi == BIG) { fail(p, "statement without ';'"); return NULL; }

	int uninit = first >= semi;
	char word[SWGL_NAME_LEN * 2];
	str_until(p, blank, 0, word, sizeof(word));
	int decl_type = type_from_str(word);

	if (decl_type == SWT_UNKNOWN)
	{
		if (!uninit && op == OP_ASSIGN)
		{
			node* lhs = parse_subexpr(p, first);
			if (!lhs) return NULL;
			p->at = second;
			node* rhs = parse_expr(p, semi);
			if (!rhs) { fail(p, "assignment without value"); return NULL; }
			node* as = new_node(p, N_ASSIGN);
			if (!as) return NULL;
			as->op = OP_ASSIGN;
			as->a = lhs;
			as->b = rhs;
			return as;
		}
		node* e = parse_expr(p, semi);
		if (!e) fail(p, "empty statement");
		return e;
	}

	if (!uninit && op != OP_ASSIGN) { fail(p, "declaration followed by a non-assignment operator"); return NULL; }

	p->at = blank + 1;
	char name[SWGL_NAME_LEN * 2];
	str_until(p, first < semi ? first : semi, 1, name, sizeof(name));
	int var = add_var(p, name, decl_type, 1);
	if (var < 0) return NULL;
	p->scope_last = p->sh->n_vars;

	if (first > semi) { p->at = semi + 1; return NULL; }

	node* d = new_node(p, N_DECL);
	if (!d) return NULL;
	d->var = var;
	p->at = second;
	d->b = parse_expr(p, semi);
	if (!d->b) { fail(p, "declaration without value"); return NULL; }
	return d;
}

static void parse_function(parser* p) /* swgl.c:1490-1576 */
{
	skip_blanks(p);
	int blank = tell_next(p, ' ');
	if (blank == BIG) { fail(p, "unexpected end of shader"); return; }
	p->at = blank + 1;
	skip_blanks(p);
	int open = tell_next(p, '(');
	int close = tell_next(p, ')');
	if (open == BIG || close == BIG) { fail(p, "function header expected"); return; }
	char fname[SWGL_NAME_LEN * 2];
	str_until(p, open, 0, fname, sizeof(fname));
	int is_main = !strcmp(fname, "main");

	p->scope_first = p->scope_last = p->sh->n_vars;
	p->at = open + 1;
	int guard = 0;
	while (p->at < close)
	{
		if (++guard > 32) { fail(p, "parameter list does not terminate"); return; }
		skip_blanks(p);
		if (p->at >= close) break;
		int b = tell_next(p, ' ');
		char tname[SWGL_NAME_LEN * 2];
		str_until(p, b, 0, tname, sizeof(tname));
		int ptype = type_from_str(tname);
		if (ptype == SWT_UNKNOWN) { fail(p, "unknown parameter type"); return; }
		p->at = b + 1;
		skip_blanks(p);
		int pend = tell_next(p, ',');
		if (pend > close) pend = close;
		char pname[SWGL_NAME_LEN * 2];
		str_until(p, pend, 1, pname, sizeof(pname));
		if (add_var(p, pname, ptype, 1) < 0) return;
		p->scope_last = p->sh->n_vars;
		p->at = pend + 1;
	}
	p->at = close + 1;
	skip_blanks(p);
	if (ch(p, p->at) != '{') { fail(p, "'{' expected"); return; }
	int block_end = tell_next_matching(p, '{', '}');
	if (block_end == BIG) { fail(p, "unbalanced '{'"); return; }
	p->at++;

	guard = 0;
	while (p->at < block_end && !p->failed)
	{
		if (++guard > MAX_LINES) { fail(p, "function body does not terminate"); return; }
		int save = p->at;
		skip_blanks(p);
		if (p->at >= block_end) break; /* the reference would run off the end here */
		p->at = save;
		node* line = parse_line(p);
		if (p->failed) return;
		if (line && is_main)
		{
			if (p->n_main_lines >= MAX_LINES) { fail(p, "too many statements"); return; }
			p->main_lines[p->n_main_lines++] = line;
		}
		if (p->at <= save) { fail(p, "statement does not advance"); return; }
	}
	if (is_main) p->sh->has_main = 1;
	p->at = block_end + 1;
	skip_blanks(p);
}

/* `uniform|in|out T name;` (swgl.c:1578-1707) */
static void parse_qualified(parser* p, int kind)
{
	skip_blanks(p);
	int semi = tell_next(p, ';');
	int blank = tell_next(p, ' ');
	if (blank > semi || blank == BIG) return;
	char tname[SWGL_NAME_LEN * 2];
	str_until(p, blank, 0, tname, sizeof(tname));
	int type = type_from_str(tname);
	if (type == SWT_UNKNOWN) return;
	p->at = blank + 1;
	skip_blanks(p);
	char name[SWGL_NAME_LEN * 2];
	str_until(p, semi, 1, name, sizeof(name));
	int v = add_var(p, name, type, 0);
	if (v < 0) return;
	if (kind == 0) p->sh->vars[v].is_uniform = 1;
	if (kind == 1) p->sh->vars[v].is_in = 1;
	if (kind == 2) p->sh->vars[v].is_out = 1;
	p->at = semi + 1;
}

/* `layout (location = N) T name;` -- cursor is just after the '(' (swgl.c:1709-1786).
 * On any early return the cursor stays where the reference leaves it, so e.g.
 * `layout (location = 0) in vec4 aPos;` falls through to an `in` declaration that is never
 * fed with attribute data, exactly as in the reference. */
static void parse_layout(parser* p)
{
	int semi = tell_next(p, ';');
	int eq = tell_next(p, '=');
	int close = tell_next(p, ')');
	if (close > semi) return;
	if (eq > close) return;
	char key[SWGL_NAME_LEN * 2];
	str_until(p, eq, 1, key, sizeof(key));
	if (strcmp(key, "location")) return;
	p->at = eq + 1;
	skip_blanks(p);
	if (!is_digit_or_minus(ch(p, p->at))) return;
	char num[64];
	str_until(p, close, 1, num, sizeof(num));
	int location = swgl_glsl_atoi(num);
	p->at = close + 1;
	skip_blanks(p);
	int blank = tell_next(p, ' ');
	if (blank > semi) return;
	char tname[SWGL_NAME_LEN * 2];
	str_until(p, blank, 1, tname, sizeof(tname));
	int type = type_from_str(tname);
	if (type == SWT_UNKNOWN) return;
	p->at = blank + 1;
	skip_blanks(p);
	char name[SWGL_NAME_LEN * 2];
	str_until(p, semi, 1, name, sizeof(name));
	int v = add_var(p, name, type, 0);
	if (v < 0) return;
	p->sh->vars[v].is_layout = 1;
	p->sh->vars[v].location = location;
	p->at = semi + 1;
}

static void parse_top(parser* p) /* swgl.c:1788-1827 */
{
	skip_blanks(p);
	int blank = tell_next(p, ' ');
	int open = tell_next(p, '(');
	char bword[SWGL_NAME_LEN * 2], pword[SWGL_NAME_LEN * 2];
	str_until(p, blank, 1, bword, sizeof(bword));
	str_until(p, open, 1, pword, sizeof(pword));
	if (!strcmp(bword, "uniform")) { p->at = blank + 1; parse_qualified(p, 0); return; }
	if (!strcmp(bword, "in")) { p->at = blank + 1; parse_qualified(p, 1); return; }
	if (!strcmp(bword, "out")) { p->at = blank + 1; parse_qualified(p, 2); return; }
	if (!strcmp(pword, "layout")) { p->at = open + 1; parse_layout(p); return; }
	parse_function(p);
}

/* ------------------------------------------------------------------------------------------
 * stage 2: typing + IR emission (rules of ExecuteGLSLToken, swgl.c:2177-2840)
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
	int type;   /* SWT_* */
	int reg;    /* index into the vector-temp file, or the matrix-temp file when swt_is_mat(type) */
} value;

typedef struct
{
	parser* p;
	swgl_ir_code* code;
	int tsp, msp;
} emitter;

static void emit(emitter* e, int op, int n, int dst, int a, int b, uint32_t imm, uint32_t imm2)
{
	if (e->code->n_ops >= SWGL_MAX_OPS) { fail(e->p, "shader has too many operations"); return; }
	swgl_ir_op* o = &e->code->ops[e->code->n_ops++];
	o->op = (uint8_t)op; o->n = (uint8_t)n; o->dst = (uint16_t)dst; o->a = (uint16_t)a; o->b = (uint16_t)b;
	o->imm = imm; o->imm2 = imm2;
}

static int talloc(emitter* e)
{
	if (e->tsp >= SWGL_MAX_TEMPS) { fail(e->p, "expression too deep"); return 0; }
	int r = e->tsp++;
	if ((uint32_t)e->tsp > e->code->n_temps) e->code->n_temps = (uint32_t)e->tsp;
	return r;
}
static int mtalloc(emitter* e)
{
	if (e->msp >= SWGL_MAX_MTEMPS) { fail(e->p, "matrix expression too deep"); return 0; }
	int r = e->msp++;
	if ((uint32_t)e->msp > e->code->n_mtemps) e->code->n_mtemps = (uint32_t)e->msp;
	return r;
}

static int vec_comps(int t) { return t == SWT_FLOAT ? 1 : t == SWT_VEC2 ? 2 : t == SWT_VEC3 ? 3 : t == SWT_VEC4 ? 4 : 0; }

/* A matrix-typed value used where only x,y,z,w,i matter contributes zeros. */
static int as_vec_reg(emitter* e, value v)
{
	if (!swt_is_mat(v.type)) return v.reg;
	int r = talloc(e);
	emit(e, SWOP_ZERO, 0, r, 0, 0, 0, 0);
	return r;
}

static value unknown_value(emitter* e, int base_t, int base_m)
{
	e->tsp = base_t; e->msp = base_m;
	value r = { SWT_UNKNOWN, talloc(e) };
	emit(e, SWOP_ZERO, 0, r.reg, 0, 0, 0, 0);
	return r;
}

static value compile_node(emitter* e, node* n);

/* AssignToExVal (swgl.c:1894-1974): silently ignored unless the types agree (sampler <- int ok) */
static void emit_store(emitter* e, int var, value v)
{
	const swgl_var* dst = &e->p->sh->vars[var];
	if (dst->type != v.type && !(dst->type == SWT_SAMPLER2D && v.type == SWT_INT)) return;
	if (vec_comps(v.type)) emit(e, SWOP_STV, vec_comps(v.type), (int)dst->word, v.reg, 0, 0, 0);
	else if (v.type == SWT_INT) emit(e, SWOP_STI, 1, (int)dst->word, v.reg, 0, 0, 0);
	else if (swt_is_mat(v.type)) emit(e, SWOP_STM, swt_mat_dim(v.type), (int)dst->word, v.reg, 0, 0, 0);
	/* SAMPLER2D / UNKNOWN values have no store case in the reference */
}

static value compile_node(emitter* e, node* n)
{
	parser* p = e->p;
	const int base_t = e->tsp, base_m = e->msp;
	value r = { SWT_UNKNOWN, 0 };
	if (!n || p->failed) { fail(p, "internal: missing node"); return r; }

	switch (n->kind)
	{
	case N_VAR:
	{
		const swgl_var* v = &p->sh->vars[n->var];
		r.type = v->type;
		if (vec_comps(v->type)) { r.reg = talloc(e); emit(e, SWOP_LDV, vec_comps(v->type), r.reg, (int)v->word, 0, 0, 0); }
		else if (v->type == SWT_INT || v->type == SWT_SAMPLER2D) { r.reg = talloc(e); emit(e, SWOP_LDI, 1, r.reg, (int)v->word, 0, 0, 0); }
		else if (swt_is_mat(v->type)) { r.reg = mtalloc(e); emit(e, SWOP_LDM, swt_mat_dim(v->type), r.reg, (int)v->word, 0, 0, 0); }
		else fail(p, "variable of unknown type");
		return r;
	}
	case N_CONST:
	{
		r.reg = talloc(e);
		if (n->is_float)
		{
			uint32_t bits; memcpy(&bits, &n->f, 4);
			r.type = SWT_FLOAT; emit(e, SWOP_CONF, 1, r.reg, 0, 0, bits, 0);
		}
		else { r.type = SWT_INT; emit(e, SWOP_CONI, 1, r.reg, 0, 0, (uint32_t)n->i, 0); }
		return r;
	}
	case N_DECL:
	{
		value v = compile_node(e, n->b);
		emit_store(e, n->var, v);
		return unknown_value(e, base_t, base_m); /* `{ GLSL_UNKNOWN }` (swgl.c:2260) */
	}
	case N_ASSIGN:
	{
		if (!n->a || n->a->kind != N_VAR) { fail(p, "assignment target must be a plain variable"); return r; }
		value v = compile_node(e, n->b);
		emit_store(e, n->a->var, v);
		return v; /* the assignment yields the assigned value (swgl.c:2270) */
	}
	case N_BIN:
	{
		if (n->op == OP_LT || n->op == OP_GT || n->op == OP_EQ)
		{
			fail(p, "comparison operators are tokenised but not executable in the reference");
			return r;
		}
		value a = compile_node(e, n->a);
		value b = compile_node(e, n->b);
		if (p->failed) return r;
		if (n->op == OP_MUL && swt_is_mat(a.type))
		{
			int dim = swt_mat_dim(a.type);
			if (b.type == a.type)
			{
				e->tsp = base_t; e->msp = base_m;
				r.type = a.type; r.reg = mtalloc(e);
				emit(e, SWOP_MULMM, dim, r.reg, a.reg, b.reg, 0, 0);
				return r;
			}
			if (vec_comps(b.type) == dim)
			{
				e->tsp = base_t; e->msp = base_m;
				r.type = b.type; r.reg = talloc(e);
				emit(e, SWOP_MULMV, dim, r.reg, a.reg, b.reg, 0, 0);
				return r;
			}
			/* any other right operand: the left matrix passes through unchanged (swgl.c:2418-2462) */
			e->tsp = base_t; e->msp = base_m;
			r.type = a.type; r.reg = mtalloc(e);
			return r;
		}
		if (a.type != b.type) return unknown_value(e, base_t, base_m);
		if (swt_is_mat(a.type))
		{
			e->tsp = base_t; e->msp = base_m;
			r.type = a.type; r.reg = mtalloc(e);
			if (n->op == OP_ADD) emit(e, SWOP_ADDM, swt_mat_dim(a.type), r.reg, a.reg, b.reg, 0, 0);
			else if (n->op == OP_SUB) emit(e, SWOP_SUBM, swt_mat_dim(a.type), r.reg, a.reg, b.reg, 0, 0);
			/* DIV leaves the matrix fields untouched (swgl.c:2475-2479) */
			return r;
		}
		e->tsp = base_t; e->msp = base_m;
		r.type = a.type; r.reg = talloc(e);
		int op = n->op == OP_ADD ? SWOP_ADD : n->op == OP_SUB ? SWOP_SUB : n->op == OP_MUL ? SWOP_MUL : SWOP_DIV;
		emit(e, op, 4, r.reg, a.reg, b.reg, 0, 0);
		return r;
	}
	case N_SWZ:
	{
		value a = compile_node(e, n->a);
		if (p->failed) return r;
		if (n->n_swz == 0) return unknown_value(e, base_t, base_m);
		if (n->n_swz > 4) { fail(p, "swizzle with more than four components"); return r; }
		int src = as_vec_reg(e, a);
		uint32_t pat = 0;
		for (int i = 0; i < n->n_swz; i++) pat |= (uint32_t)n->swz[i] << (2 * i);
		e->tsp = base_t; e->msp = base_m;
		r.type = n->n_swz == 1 ? SWT_FLOAT : n->n_swz == 2 ? SWT_VEC2 : n->n_swz == 3 ? SWT_VEC3 : SWT_VEC4;
		r.reg = talloc(e);
		emit(e, SWOP_SWZ, n->n_swz, r.reg, src, 0, pat, 0);
		return r;
	}
	case N_CALL:
	{
		value args[8];
		int regs[8];
		int na = n->n_args > 8 ? 8 : n->n_args;
		int need = (n->op == FN_COS || n->op == FN_SIN || n->op == FN_TAN) ? 1 : 2;
		if (n->n_args != need) return unknown_value(e, base_t, base_m); /* Args.Size check comes first */
		for (int i = 0; i < na; i++) args[i] = compile_node(e, n->args[i]);
		if (p->failed) return r;
		if (n->op == FN_TEXTURE)
		{
			if (args[0].type != SWT_SAMPLER2D || args[1].type != SWT_VEC2) return unknown_value(e, base_t, base_m);
			e->tsp = base_t; e->msp = base_m;
			r.type = SWT_VEC4; r.reg = talloc(e);
			emit(e, SWOP_TEX, 4, r.reg, args[0].reg, args[1].reg, 0, 0);
			return r;
		}
		if (swt_is_mat(args[0].type))
		{
			/* xyzw of a matrix value are not observable afterwards: the matrix passes through */
			e->tsp = base_t; e->msp = base_m;
			r.type = args[0].type; r.reg = mtalloc(e);
			return r;
		}
		for (int i = 0; i < na; i++) regs[i] = as_vec_reg(e, args[i]);
		e->tsp = base_t; e->msp = base_m;
		r.type = args[0].type; r.reg = talloc(e);
		int op = n->op == FN_COS ? SWOP_COS : n->op == FN_SIN ? SWOP_SIN : n->op == FN_TAN ? SWOP_TAN
		       : n->op == FN_MIN ? SWOP_MIN : SWOP_MAX;
		emit(e, op, 4, r.reg, regs[0], need == 2 ? regs[1] : 0, 0, 0);
		return r;
	}
	case N_CONS:
	{
		int regs[4] = { 0, 0, 0, 0 };
		uint32_t int_mask = 0;
		for (int i = 0; i < n->n_args; i++)
		{
			value a = compile_node(e, n->args[i]);
			if (p->failed) return r;
			if (a.type == SWT_INT) int_mask |= 1u << i;
			regs[i] = as_vec_reg(e, a);
		}
		e->tsp = base_t; e->msp = base_m;
		r.reg = talloc(e);
		if (n->op == SWT_INT)
		{
			r.type = SWT_INT;
			emit(e, SWOP_ICONS, 1, r.reg, regs[0], 0, 0, int_mask);
		}
		else
		{
			r.type = n->op;
			emit(e, SWOP_CONS, n->n_args, r.reg, regs[0], regs[1],
			     (uint32_t)regs[2] | ((uint32_t)regs[3] << 16), int_mask);
		}
		return r;
	}
	}
	fail(p, "internal: unknown node");
	return r;
}

/* ---- shape recognition for the specialised device functors ---- */
static void recognise(parser* p)
{
	swgl_shader* sh = p->sh;
	sh->n_statements = p->n_main_lines;
	sh->simple_copies = 1;
	sh->pos_kind = 0;
	sh->pos_attr_var = sh->pos_mat_var = -1;
	sh->fs_kind = SWFS_GENERIC;
	sh->fs_in_var = sh->fs_sampler_var = -1;
	int assigned[SWGL_MAX_VARS];
	memset(assigned, 0, sizeof(assigned));

	for (int i = 0; i < p->n_main_lines; i++)
	{
		node* n = p->main_lines[i];
		if (n->kind != N_ASSIGN || !n->a || n->a->kind != N_VAR) { sh->simple_copies = 0; continue; }
		int dst = n->a->var;
		if (assigned[dst]++) sh->simple_copies = 0;
		node* rhs = n->b;
		const swgl_var* dv = &sh->vars[dst];
		if (rhs->kind == N_VAR && sh->vars[rhs->var].type == dv->type && vec_comps(dv->type))
		{
			sh->vars[dst].copy_of = rhs->var;
			if (dst == 0) { sh->pos_kind = 1; sh->pos_attr_var = rhs->var; }
			continue;
		}
		if (dst == 0 && rhs->kind == N_BIN && rhs->op == OP_MUL && rhs->a->kind == N_VAR && rhs->b->kind == N_VAR
		    && sh->vars[rhs->a->var].type == SWT_MAT4 && sh->vars[rhs->a->var].is_uniform
		    && sh->vars[rhs->b->var].type == SWT_VEC4)
		{
			sh->pos_kind = 2; sh->pos_mat_var = rhs->a->var; sh->pos_attr_var = rhs->b->var;
			continue;
		}
		sh->simple_copies = 0;
	}
	/* a copy is only "simple" if its source is never written by main() */
	for (int v = 0; v < sh->n_vars; v++)
		if (sh->vars[v].copy_of >= 0 && assigned[sh->vars[v].copy_of]) sh->simple_copies = 0;
	if (sh->pos_attr_var >= 0 && assigned[sh->pos_attr_var]) sh->simple_copies = 0;

	/* fragment shapes: exactly one statement, writing the first `out` variable */
	if (p->n_main_lines == 1)
	{
		node* n = p->main_lines[0];
		int first_out = -1;
		for (int v = 0; v < sh->n_globals; v++) if (sh->vars[v].is_out) { first_out = v; break; }
		if (n->kind == N_ASSIGN && n->a && n->a->kind == N_VAR && n->a->var == first_out
		    && first_out >= 0 && sh->vars[first_out].type == SWT_VEC4)
		{
			node* rhs = n->b;
			if (rhs->kind == N_VAR && sh->vars[rhs->var].is_in && sh->vars[rhs->var].type == SWT_VEC4)
			{
				sh->fs_kind = SWFS_VARYING; sh->fs_in_var = rhs->var;
			}
			else if (rhs->kind == N_CALL && rhs->op == FN_TEXTURE && rhs->n_args == 2
			         && rhs->args[0]->kind == N_VAR && sh->vars[rhs->args[0]->var].type == SWT_SAMPLER2D
			         && sh->vars[rhs->args[0]->var].is_uniform)
			{
				node* uv = rhs->args[1];
				if (uv->kind == N_VAR && sh->vars[uv->var].is_in && sh->vars[uv->var].type == SWT_VEC2)
				{
					sh->fs_kind = SWFS_TEXTURE; sh->fs_in_var = uv->var; sh->fs_sampler_var = rhs->args[0]->var;
					sh->fs_swz[0] = 0; sh->fs_swz[1] = 1;
				}
				else if (uv->kind == N_SWZ && uv->n_swz == 2 && uv->a->kind == N_VAR && sh->vars[uv->a->var].is_in
				         && vec_comps(sh->vars[uv->a->var].type) > uv->swz[0]
				         && vec_comps(sh->vars[uv->a->var].type) > uv->swz[1])
				{
					sh->fs_kind = SWFS_TEXTURE; sh->fs_in_var = uv->a->var; sh->fs_sampler_var = rhs->args[0]->var;
					sh->fs_swz[0] = uv->swz[0]; sh->fs_swz[1] = uv->swz[1];
				}
			}
		}
	}
}

/* ------------------------------------------------------------------------------------------
 * entry points
 * ---------------------------------------------------------------------------------------- */
static uint64_t g_next_shader_id = 1;

swgl_shader* swgl_glsl_compile(const char* source)
{
	swgl_shader* sh = (swgl_shader*)calloc(1, sizeof(swgl_shader));
	parser* p = (parser*)calloc(1, sizeof(parser));
	if (!sh || !p) { free(sh); free(p); return NULL; }
	sh->id = g_next_shader_id++;
	p->sh = sh;

	/* swgl.c:1839-1843: newlines and tabs are deleted, no blank is inserted */
	size_t len = strlen(source);
	if (len >= (size_t)BIG - 16) { free(sh); free(p); return NULL; }
	p->code = (char*)calloc(len + 16, 1);
	if (!p->code) { free(sh); free(p); return NULL; }
	for (size_t i = 0; i < len; i++)
		if (source[i] != '\n' && source[i] != '\t') p->code[p->size++] = source[i];

	add_var(p, "gl_Position", SWT_VEC4, 0); /* pre-declared (swgl.c:1844-1854) */

	int guard = 0;
	while (p->at < p->size - 2 && !p->failed) /* swgl.c:1856 */
	{
		int save = p->at;
		parse_top(p);
		if (p->at <= save) { fail(p, "top-level declaration does not advance"); break; }
		if (++guard > 512) { fail(p, "shader does not terminate"); break; }
	}

	if (!p->failed && !sh->has_main) fail(p, "no main()");
	if (!p->failed)
	{
		emitter e = { p, &sh->code, 0, 0 };
		for (int i = 0; i < p->n_main_lines && !p->failed; i++)
		{
			e.tsp = 0; e.msp = 0;
			compile_node(&e, p->main_lines[i]);
		}
		if (sh->n_vars > 0)
		{
			const swgl_var* last = &sh->vars[sh->n_vars - 1];
			sh->code.n_words = last->word + (uint32_t)swt_words(last->type);
		}
	}
	if (!p->failed) recognise(p);
	sh->ok = !p->failed;
	free(p->code);
	free(p);
	return sh;
}

void swgl_glsl_free(swgl_shader* s) { free(s); }

int swgl_glsl_find_var(const swgl_shader* s, const char* name)
{
	for (int i = 0; i < s->n_globals; i++)
		if (!s->vars[i].is_local && !strcmp(s->vars[i].name, name)) return i;
	return -1;
}

static const char* type_name(int t)
{
	static const char* names[] = { "float", "vec2", "vec3", "vec4", "int", "mat2", "mat3", "mat4", "sampler2D", "unknown" };
	return (t >= 0 && t <= SWT_UNKNOWN) ? names[t] : "?";
}

size_t swgl_glsl_dump(const swgl_shader* s, char* buf, size_t len)
{
	static const char* op_names[SWOP__COUNT] = {
		"nop", "ldv", "ldi", "ldm", "conf", "coni", "zero", "stv", "sti", "stm", "add", "sub", "mul", "div",
		"addm", "subm", "mulmm", "mulmv", "tex", "sin", "cos", "tan", "min", "max", "swz", "cons", "icons", "movm2t" };
	size_t need = 0;
#define OUT(...) do { int k_ = snprintf(buf && need < len ? buf + need : NULL, buf && need < len ? len - need : 0, __VA_ARGS__); if (k_ > 0) need += (size_t)k_; } while (0)
	OUT("ok=%d", s->ok);
	if (!s->ok) OUT(" error=\"%s\"", s->error);
	OUT("\nshape: statements=%d simple_copies=%d pos_kind=%d fs_kind=%d\n", s->n_statements, s->simple_copies, s->pos_kind, s->fs_kind);
	for (int i = 0; i < s->n_vars; i++)
	{
		const swgl_var* v = &s->vars[i];
		OUT("var %d %s %s word=%u%s%s%s%s%s", i, type_name(v->type), v->name, v->word,
		    v->is_uniform ? " uniform" : "", v->is_in ? " in" : "", v->is_out ? " out" : "",
		    v->is_local ? " local" : "", v->is_layout ? " layout" : "");
		if (v->is_layout) OUT("(location=%d)", v->location);
		if (v->copy_of >= 0) OUT(" copy_of=%d", v->copy_of);
		OUT("\n");
	}
	for (uint32_t i = 0; i < s->code.n_ops; i++)
	{
		const swgl_ir_op* o = &s->code.ops[i];
		OUT("op %u %s n=%u dst=%u a=%u b=%u imm=0x%x imm2=0x%x\n", i, op_names[o->op], o->n, o->dst, o->a, o->b, o->imm, o->imm2);
	}
#undef OUT
	if (buf && len) buf[need < len ? need : len - 1] = 0;
	return need;
}
