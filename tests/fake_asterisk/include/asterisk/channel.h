/* fake <asterisk/channel.h> (test infrastructure): a scripted channel that delivers 20 ms slinear frames */
#ifndef FAKE_AST_CHANNEL_H_
#define FAKE_AST_CHANNEL_H_
#include <sys/time.h>
struct ast_channel;
enum ast_channel_state { AST_STATE_DOWN, AST_STATE_RESERVED, AST_STATE_OFFHOOK, AST_STATE_DIALING, AST_STATE_RING, AST_STATE_RINGING, AST_STATE_UP, AST_STATE_BUSY };
enum ast_frame_type { AST_FRAME_DTMF_END = 1, AST_FRAME_VOICE, AST_FRAME_VIDEO, AST_FRAME_CONTROL, AST_FRAME_NULL };
struct ast_frame {
  enum ast_frame_type frametype;
  int datalen, samples;
  struct { void *ptr; } data;
};
enum ast_channel_state ast_channel_state(const struct ast_channel *chan);
int ast_answer(struct ast_channel *chan);
int ast_waitfor(struct ast_channel *chan, int ms);
struct ast_frame *ast_read(struct ast_channel *chan);
void ast_frfree(struct ast_frame *fr);
struct timeval ast_tvnow(void);
int ast_remaining_ms(struct timeval start, int max_ms);
#endif
