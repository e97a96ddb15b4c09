/*
 * rlfc_client.c -- the CFD client process of RLFluidControl, rebuilt on librlfc.so (C, no Processing/JVM).
 *
 * It replaces clientLilypad/clientCFD.pde and speaks the reference's unchanged wire protocol to the
 * unchanged agent server (server/server.py:31,51-59): XML-RPC over HTTP POST to http://localhost:8000.
 * Call order and payloads follow the sketch exactly:
 *   setup()            init(int -1)                                         clientCFD.pde:18-33
 *   setUpNewSim()      new env resumed from init.bdim; start_episode(int -1) clientCFD.pde:93-114
 *   draw() while t<50  update2(); log line; after t>1: callLearn--, Cd+=fx, Cl+=fy; every 16th step
 *                      Cd=Cd/16*2/24, Cl=..., request_stochastic_action("<Cl>_<Cd>") -> "<a1>_<a2>",
 *                      xi = a, xi_m = 5*xi                                   clientCFD.pde:35-55,117-133
 *   at t>=50           train(int 1000); save(int 1000); next episode          clientCFD.pde:66-84
 * Quirks kept: callLearn/Cd/Cl are sketch globals that survive episodes and are never zeroed
 * (clientCFD.pde:11-13,44-47); a failed RPC leaves the action at (0,0) (clientCFD.pde:118-132).
 * Per-step trace files use the SaveScalar format (SaveScalar.pde:28-33,61-72).
 *
 * The protocol is single-environment (server.py:157-165 pairs consecutive records as transitions), so this
 * faithful driver runs one environment per server; batched training uses rlfc_env_step from a host program.
 *
 * usage: rlfc_client [--host H] [--port P] [--episodes N] [--time T] [--init-time T0] [--init FILE]
 *                    [--save-dir DIR] [--train-steps K] [--max-steps S] [--deterministic] [--quiet]
 */
#define _POSIX_C_SOURCE 200809L
#include <arpa/inet.h>
#include <errno.h>
#include <netdb.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/rlfc.h"

/* ------------------------------------------------------------------ XML-RPC over HTTP/1.0 ---- */
static int http_post(const char *host, int port, const char *body, char *resp, size_t cap) {
  struct addrinfo hints, *res = NULL;
  char ports[16];
  memset(&hints, 0, sizeof hints);
  hints.ai_family = AF_UNSPEC;
  hints.ai_socktype = SOCK_STREAM;
  snprintf(ports, sizeof ports, "%d", port);
  if (getaddrinfo(host, ports, &hints, &res) != 0 || !res) return -1;
  int fd = -1;
  for (struct addrinfo *a = res; a; a = a->ai_next) {
    fd = socket(a->ai_family, a->ai_socktype, a->ai_protocol);
    if (fd < 0) continue;
    if (connect(fd, a->ai_addr, a->ai_addrlen) == 0) break;
    close(fd);
    fd = -1;
  }
  freeaddrinfo(res);
  if (fd < 0) return -1;
  char hdr[512];
  int hl = snprintf(hdr, sizeof hdr,
                    "POST / HTTP/1.0\r\nHost: %s:%d\r\nUser-Agent: rlfc_client\r\nContent-Type: text/xml\r\n"
                    "Content-Length: %zu\r\n\r\n", host, port, strlen(body));
  if (write(fd, hdr, (size_t)hl) != hl || write(fd, body, strlen(body)) != (ssize_t)strlen(body)) { close(fd); return -1; }
  size_t got = 0;
  for (;;) {
    ssize_t k = read(fd, resp + got, cap - 1 - got);
    if (k <= 0) break;
    got += (size_t)k;
    if (got >= cap - 1) break;
  }
  resp[got] = 0;
  close(fd);
  return got > 0 ? 0 : -1;
}

/* one call with a single <int> or <string> parameter; the reply's scalar payload is copied to `out` */
static int rpc_call(const char *host, int port, const char *method, const char *type, const char *value, char *out,
                    size_t outcap) {
  char body[1024], resp[8192];
  snprintf(body, sizeof body,
           "<?xml version=\"1.0\"?><methodCall><methodName>%s</methodName><params><param><value><%s>%s</%s></value>"
           "</param></params></methodCall>", method, type, value, type);
  if (http_post(host, port, body, resp, sizeof resp) != 0) return -1;
  if (strstr(resp, "<fault>")) return -2;
  const char *v = strstr(resp, "<value>");
  if (!v) return -3;
  v += 7;
  const char *s = v;
  if (*v == '<') {               /* typed: <string>..</string>, <boolean>1</boolean>, <int>.. */
    s = strchr(v, '>');
    if (!s) return -3;
    s++;
  }
  const char *e = strchr(s, '<');
  if (!e) return -3;
  size_t n = (size_t)(e - s);
  if (n >= outcap) n = outcap - 1;
  memcpy(out, s, n);
  out[n] = 0;
  return 0;
}

/* ------------------------------------------------------------------ the sketch ---- */
typedef struct {
  const char *host; int port; int episodes; float Time, initTime; const char *init; const char *save_dir;
  int train_steps; long max_steps; int deterministic; int quiet;
} options;

int rlfc_format_float_java(float v, char *buf, int cap);   /* Float.toString, exported by librlfc.so */

int main(int argc, char **argv) {
  options o = {"localhost", 8000, 1, 50.f, 1.f, NULL, "saved", 1000, -1, 0, 0};
  for (int i = 1; i < argc; i++) {
    const char *a = argv[i];
    const char *v = (i + 1 < argc) ? argv[i + 1] : NULL;
    if (!strcmp(a, "--host") && v) { o.host = v; i++; }
    else if (!strcmp(a, "--port") && v) { o.port = atoi(v); i++; }
    else if (!strcmp(a, "--episodes") && v) { o.episodes = atoi(v); i++; }
    else if (!strcmp(a, "--time") && v) { o.Time = strtof(v, NULL); i++; }
    else if (!strcmp(a, "--init-time") && v) { o.initTime = strtof(v, NULL); i++; }
    else if (!strcmp(a, "--init") && v) { o.init = v; i++; }
    else if (!strcmp(a, "--save-dir") && v) { o.save_dir = v; i++; }
    else if (!strcmp(a, "--train-steps") && v) { o.train_steps = atoi(v); i++; }
    else if (!strcmp(a, "--max-steps") && v) { o.max_steps = atol(v); i++; }
    else if (!strcmp(a, "--deterministic")) o.deterministic = 1;
    else if (!strcmp(a, "--quiet")) o.quiet = 1;
    else { fprintf(stderr, "unknown option %s\n", a); return 2; }
  }
  char reply[256], num[64];

  /* setup(): client.execute("init", -1)   clientCFD.pde:20-29 */
  if (rpc_call(o.host, o.port, "init", "int", "-1", reply, sizeof reply) != 0)
    fprintf(stderr, "init: RPC failed (%s)\n", strerror(errno));
  else if (!o.quiet) printf("init -> %s\n", reply);

  rlfc_config cfg;
  rlfc_default_config(&cfg);
  cfg.n_envs = 1;
  cfg.init_bdim_path = o.init;          /* saved/init/init.bdim of the reference, or its .bdimb form */
  cfg.init_time = o.initTime;
  cfg.episode_time = o.Time;
  rlfc_env *env = NULL;
  if (rlfc_env_create(&cfg, &env) != 0) { fprintf(stderr, "rlfc_env_create: %s\n", rlfc_last_error()); return 1; }

  /* sketch globals clientCFD.pde:9-14 */
  int callLearn = 16, simNum = 1;
  float Cd = 0, Cl = 0;
  const int resolution = cfg.resolution;
  long total_steps = 0;
  mkdir(o.save_dir, 0777);

  for (int ep = 0; ep < o.episodes; ep++, simNum++) {
    /* setUpNewSim(simNum) clientCFD.pde:93-114 */
    if (rlfc_env_reset(env, NULL, 1, 0) != 0) { fprintf(stderr, "reset: %s\n", rlfc_last_error()); return 1; }
    float xi[2] = {0.f, 0.f};
    char path[512];
    snprintf(path, sizeof path, "%s/%d.txt", o.save_dir, simNum);
    FILE *dat = fopen(path, "w");
    if (dat) fprintf(dat, "%%%% Force and pressure coefficients using processing viscous simulation\n\n"
                          "%%%% Fellowing: t, force.x, force.y\n\n");          /* SaveScalar.pde:29-33 */
    if (rpc_call(o.host, o.port, "start_episode", "int", "-1", reply, sizeof reply) != 0)
      fprintf(stderr, "start_episode: RPC failed\n");

    float t = 0.f;
    while (t < o.Time && (o.max_steps < 0 || total_steps < o.max_steps)) {      /* draw() clientCFD.pde:35-55 */
      float force[2], probes[RLFC_NUM_PROBES];
      if (rlfc_env_substep(env, xi, force, probes) != 0) { fprintf(stderr, "substep: %s\n", rlfc_last_error()); return 1; }
      rlfc_env_get_time(env, &t);
      total_steps++;
      if (dat) {                                                               /* SaveScalar.addData03 */
        const float vals[5] = {t, force[0], force[1], xi[0], xi[1]};
        for (int k = 0; k < 5; k++) { rlfc_format_float_java(vals[k], num, sizeof num); fprintf(dat, "%s ", num); }
        for (int k = 0; k < RLFC_NUM_PROBES; k++) { rlfc_format_float_java(probes[k], num, sizeof num); fprintf(dat, "%s ", num); }
        fprintf(dat, "\n");
      }
      if (t > o.initTime) {
        callLearn--;
        Cd += force[0];
        Cl += force[1];
        if (callLearn <= 0) {
          callLearn = 16;
          Cd = Cd / callLearn * 2 / resolution;
          Cl = Cl / callLearn * 2 / resolution;
          /* callAction(Cl, Cd) clientCFD.pde:117-133: String.valueOf(Cl) + "_" + String.valueOf(Cd) */
          char msg[160], a[64], b[64];
          rlfc_format_float_java(Cl, a, sizeof a);
          rlfc_format_float_java(Cd, b, sizeof b);
          snprintf(msg, sizeof msg, "%s_%s", a, b);
          float XI[2] = {0.f, 0.f};                                            /* zeros on failure */
          if (rpc_call(o.host, o.port, o.deterministic ? "request_deterministic_action" : "request_stochastic_action",
                       "string", msg, reply, sizeof reply) == 0) {
            char *us = strchr(reply, '_');
            if (us) { *us = 0; XI[0] = strtof(reply, NULL); XI[1] = strtof(us + 1, NULL); }
          } else {
            fprintf(stderr, "request_action: RPC failed, using (0,0)\n");
          }
          xi[0] = XI[0];
          xi[1] = XI[1];
          if (!o.quiet) printf("t=%.4f  Cl=%s Cd=%s -> xi=(%g, %g)\n", t, a, b, xi[0], xi[1]);
        }
      }
    }
    if (dat) fclose(dat);                                                      /* dat.finish() */
    /* episode end clientCFD.pde:72-81 */
    snprintf(num, sizeof num, "%d", o.train_steps);
    if (rpc_call(o.host, o.port, "train", "int", num, reply, sizeof reply) != 0) fprintf(stderr, "train: RPC failed\n");
    else if (!o.quiet) printf("train -> %s\n", reply);
    if (rpc_call(o.host, o.port, "save", "int", num, reply, sizeof reply) != 0) fprintf(stderr, "save: RPC failed\n");
    if (o.max_steps >= 0 && total_steps >= o.max_steps) break;
  }
  rlfc_env_destroy(env);
  return 0;
}
